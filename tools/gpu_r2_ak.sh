#!/bin/bash
O=gpurun_out/r2ak; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
for v in snall sn0; do echo "== $v"; RRL_LIB_PATH=$V/librrl_b200_$v.so timeout 300 python tools/debug_degenerate.py 2>&1 | grep "supers 1" ; done
