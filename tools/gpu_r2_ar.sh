#!/bin/bash
O=gpurun_out/r2ar; mkdir -p $O
timeout 300 python tools/stages.py large large2 large8 big > $O/stages.log 2>&1; cut -c1-110 $O/stages.log
echo "== big at waves 1 / 3 / 6" ; for w in 1 3 6; do timeout 100 python tools/stages.py big 9=$w | grep -v peak | cut -c1-110; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -m gpu -q -x > $O/tests.log 2>&1; tail -2 $O/tests.log
