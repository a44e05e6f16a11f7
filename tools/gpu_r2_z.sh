#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_large python tools/prof_one.py large 3 > $O/ncu_large.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:node_kernel -s 2 -c 1 -o $O/prof_node_large python tools/prof_one.py large 3 > $O/ncu_node.log 2>&1
timeout 100 python tools/stages.py large >> $O/stages.log 2>&1; cat $O/stages.log
ls -la $O
