#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
timeout 300 python tools/stages.py demo dcp rpm fmr large big > $O/stages.log 2>&1; cat $O/stages.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
