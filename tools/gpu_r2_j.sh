#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
RRL_LIB_PATH=$PWD/a-robust-registration-loss_b200/build/variants/librrl_b200_counters.so timeout 300 python tools/counters.py large dcp rpm > $O/counters.log 2>&1; cat $O/counters.log
timeout 300 python -m pytest tests/test_gpu_aux.py -m gpu -q -x -k "demo" > $O/tests.log 2>&1; tail -3 $O/tests.log
