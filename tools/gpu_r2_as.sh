#!/bin/bash
O=gpurun_out/r2as; mkdir -p $O
{
for it in 8 6 4 2 0; do echo "== ball iterations $it"; timeout 200 python tools/stages.py large large8 big 8=$it; done
} > $O/stages.log 2>&1
grep -v "^peak" $O/stages.log | cut -c1-110
