#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
echo "== compressed" > $O/counters.log; RRL_LIB_PATH=$V/librrl_b200_cnt1.so timeout 200 python tools/counters.py large >> $O/counters.log 2>&1
echo "== fp32" >> $O/counters.log; RRL_LIB_PATH=$V/librrl_b200_cnt0.so timeout 200 python tools/counters.py large >> $O/counters.log 2>&1
cat $O/counters.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_large python tools/prof_one.py large 3 > $O/ncu_large.log 2>&1
ls -la $O
