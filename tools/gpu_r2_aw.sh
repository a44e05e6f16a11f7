#!/bin/bash
O=gpurun_out/r2aw; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
{
echo "== product (512 threads)"; timeout 300 python tools/stages.py demo dcp rpm fmr large
echo "== tail kernel with 1024 threads"; RRL_LIB_PATH=$V/librrl_b200_tail1024.so timeout 300 python tools/stages.py demo dcp rpm fmr large
} > $O/stages.log 2>&1; grep -v "^peak" $O/stages.log | sed 's/"prep.*"median"/.../' | cut -c1-120
