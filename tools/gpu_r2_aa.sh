#!/bin/bash
O=gpurun_out/r2aa; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
{
echo "== product"; timeout 200 python tools/stages.py large large8
echo "== LPT 1 (6=1)"; timeout 200 python tools/stages.py large large8 6=1
echo "== LPT 4 (6=4)"; timeout 200 python tools/stages.py large 6=4
echo "== target waves 8 (9=8)"; timeout 200 python tools/stages.py large large8 9=8
echo "== target waves 32 (9=32)"; timeout 200 python tools/stages.py large large8 9=32
echo "== level-1 prefetch"; RRL_LIB_PATH=$V/librrl_b200_pf1.so timeout 200 python tools/stages.py large large8
echo "== 4 CTAs/SM in super-node mode"; RRL_LIB_PATH=$V/librrl_b200_mb4.so timeout 200 python tools/stages.py large large8
} > $O/stages.log 2>&1
grep -v "^peak" $O/stages.log | cut -c1-200
