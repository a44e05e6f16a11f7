#!/bin/bash
O=gpurun_out/r2ao; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "randomised" > $O/tests_sweep.log 2>&1; tail -8 $O/tests_sweep.log
{
echo "== product"; timeout 200 python tools/stages.py dcp fmr
for m in 4 6 12 16; do echo "== exact grid $m CTAs/SM"; timeout 200 python tools/stages.py dcp fmr 13=$m; done
echo "== exact ILP 4, 16 CTAs/SM"; timeout 200 python tools/stages.py dcp 13=16 11=4
echo "== exact ILP 1, 16 CTAs/SM"; timeout 200 python tools/stages.py dcp 13=16 11=1
} > $O/stages.log 2>&1
grep -v "^peak" $O/stages.log | cut -c1-120
