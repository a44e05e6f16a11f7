#!/bin/bash
O=gpurun_out/r2am; mkdir -p $O
{
echo "== product"; timeout 200 python tools/stages.py dcp rpm
echo "== LPT 1"; timeout 200 python tools/stages.py dcp rpm 6=1
echo "== LPT 4"; timeout 200 python tools/stages.py dcp 6=4
echo "== waves 8 (2=8)"; timeout 200 python tools/stages.py dcp 2=8
echo "== waves 32 (2=32)"; timeout 200 python tools/stages.py dcp 2=32
} > $O/stages.log 2>&1
grep -v "^peak" $O/stages.log | cut -c1-170
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_large python tools/prof_one.py large 3 > $O/ncu_large.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_dcp python tools/prof_one.py dcp 3 > $O/ncu_dcp.log 2>&1
