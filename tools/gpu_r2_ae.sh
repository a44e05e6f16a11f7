#!/bin/bash
O=gpurun_out/r2ae; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_large python tools/prof_one.py large 3 > $O/ncu_large.log 2>&1
ls -la $O
