#!/bin/bash
# round 2, GPU run K (1 GPU): sector-aligned records in super-node mode; tighter guard constants (A/B)
O=gpurun_out/r2k; mkdir -p $O
V=$PWD/a-robust-registration-loss_b200/build/variants
echo "== default (aligned strides)" >> $O/stages.log; timeout 200 python tools/stages.py large big large8 >> $O/stages.log 2>&1
echo "== guard 64/40" >> $O/stages.log; RRL_LIB_PATH=$V/librrl_b200_guard.so timeout 200 python tools/stages.py large big large8 >> $O/stages.log 2>&1
cat $O/stages.log
RRL_LIB_PATH=$V/librrl_b200_counters.so timeout 200 python tools/counters.py large > $O/counters.log 2>&1
RRL_LIB_PATH=$V/librrl_b200_guardc.so timeout 200 python tools/counters.py large >> $O/counters.log 2>&1; cat $O/counters.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
RRL_LIB_PATH=$V/librrl_b200_guard.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/tests_guard.log 2>&1; tail -3 $O/tests_guard.log
