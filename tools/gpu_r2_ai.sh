#!/bin/bash
O=gpurun_out/r2ai; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
RRL_LIB_PATH=$V/librrl_b200_marks.so timeout 200 python tools/marks.py dcp demo > $O/marks.log 2>&1; grep -v "^peak" $O/marks.log
