#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
echo "== single call sites + compressed level-2 records (product)" > $O/stages.log; timeout 300 python tools/stages.py demo dcp rpm fmr large big large8 >> $O/stages.log 2>&1
echo "== single call sites, fp32 records (RRL_PC8=0)" >> $O/stages.log; RRL_LIB_PATH=$V/librrl_b200_pc0.so timeout 300 python tools/stages.py large big large8 >> $O/stages.log 2>&1
cat $O/stages.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -m gpu -q -x > $O/tests.log 2>&1; tail -5 $O/tests.log
