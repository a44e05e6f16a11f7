#!/bin/bash
O=gpurun_out/r2al; mkdir -p $O
timeout 300 python tools/stages.py demo dcp rpm fmr large big large8 > $O/stages.log 2>&1; cat $O/stages.log
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
