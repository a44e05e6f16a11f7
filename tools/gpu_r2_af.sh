#!/bin/bash
O=gpurun_out/r2af; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
{
echo "== product"; timeout 200 python tools/stages.py large large8
echo "== LPT 1 (6=1)"; timeout 200 python tools/stages.py large large8 6=1
for w in 3 4 8 12; do echo "== target waves $w"; timeout 200 python tools/stages.py large large8 9=$w; done
echo "== level-1 prefetch"; RRL_LIB_PATH=$V/librrl_b200_pf1.so timeout 200 python tools/stages.py large large8
echo "== ball iterations 4 (8=4)"; timeout 200 python tools/stages.py large 8=4
echo "== ball iterations 12 (8=12)"; timeout 200 python tools/stages.py large 8=12
} > $O/stages.log 2>&1
grep -v "^peak" $O/stages.log | cut -c1-200
