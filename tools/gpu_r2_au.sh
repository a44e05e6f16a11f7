#!/bin/bash
O=gpurun_out/r2au; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:node_kernel -s 2 -c 1 -o $O/prof_node_large python tools/prof_one.py large 3 > $O/ncu_node.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sort_keys_kernel -s 2 -c 1 -o $O/prof_keys_large python tools/prof_one.py large 3 > $O/ncu_keys.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prep_kernel -s 2 -c 1 -o $O/prof_prep_large python tools/prof_one.py large 3 > $O/ncu_prep.log 2>&1
ls -la $O
