#!/bin/bash
O=gpurun_out/r2av; mkdir -p $O
{
echo "== hand-off"; timeout 300 python tools/stages.py demo dcp rpm fmr
echo "== no hand-off (15=1)"; timeout 300 python tools/stages.py demo dcp rpm fmr 15=1
} > $O/stages.log 2>&1; grep -v "^peak" $O/stages.log | cut -c1-100
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -2 $O/tests.log
