#!/bin/bash
O=gpurun_out/r2ab; mkdir -p $O
{
echo "== product (waves 8, packed level 1, approximate centre heuristics)"; timeout 200 python tools/stages.py large large8 large2 big dcp demo
for w in 4 6 12; do echo "== target waves $w"; timeout 200 python tools/stages.py large large8 large2 9=$w; done
} > $O/stages.log 2>&1
grep -v "^peak" $O/stages.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
