#!/bin/bash
O=gpurun_out/r2aj; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
RRL_LIB_PATH=$V/librrl_b200_marks.so timeout 200 python tools/marks.py dcp demo > $O/marks.log 2>&1; grep "tail marks\|build marks\|build CTAs" $O/marks.log
timeout 300 python tools/stages.py demo dcp rpm fmr large > $O/stages.log 2>&1; cat $O/stages.log
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
