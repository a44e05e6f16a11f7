#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
RRL_LIB_PATH=$V/librrl_b200_marks.so timeout 200 python tools/marks.py dcp demo > $O/marks.log 2>&1; grep -v "^peak" $O/marks.log
echo "== product (hoisted backward loads)" > $O/stages.log; timeout 200 python tools/stages.py demo dcp >> $O/stages.log 2>&1
echo "== bwd0" >> $O/stages.log; RRL_LIB_PATH=$V/librrl_b200_bwd0.so timeout 200 python tools/stages.py demo dcp >> $O/stages.log 2>&1
echo "== product again" >> $O/stages.log; timeout 200 python tools/stages.py demo dcp >> $O/stages.log 2>&1
cat $O/stages.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
