#!/bin/bash
O=gpurun_out/r2aq; mkdir -p $O
{
for w in 1 2 3 6; do echo "== target waves $w"; timeout 200 python tools/stages.py large8 large2 large 9=$w; done
echo "== waves 2, LPT 1"; timeout 200 python tools/stages.py large8 large2 9=2 6=1
echo "== waves 1, LPT 1"; timeout 200 python tools/stages.py large8 large2 9=1 6=1
} > $O/stages.log 2>&1
grep -v "^peak" $O/stages.log | cut -c1-110
