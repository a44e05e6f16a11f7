#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
for kv in "" "10=2" "1=8" "8=4" "8=12"; do echo "== $kv" >> $O/stages.log; timeout 200 python tools/stages.py large large8 $kv >> $O/stages.log 2>&1; done; cat $O/stages.log
