#!/bin/bash
O=gpurun_out/r2ad; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
RRL_LIB_PATH=$V/librrl_b200_cnt1.so timeout 200 python tools/counters.py large dcp > $O/counters.log 2>&1; cat $O/counters.log
timeout 300 python tools/stages.py large large8 large2 big dcp rpm > $O/stages.log 2>&1; cat $O/stages.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
