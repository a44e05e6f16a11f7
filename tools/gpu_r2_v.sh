#!/bin/bash
O=gpurun_out/r2v; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
echo "== product" > $O/stages.log; timeout 300 python tools/stages.py dcp large big >> $O/stages.log 2>&1
echo "== level 2 reads half a node (wrong results, timing only)" >> $O/stages.log; RRL_LIB_PATH=$V/librrl_b200_half.so timeout 300 python tools/stages.py large big >> $O/stages.log 2>&1
cat $O/stages.log
