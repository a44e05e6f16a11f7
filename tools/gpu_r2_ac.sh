#!/bin/bash
O=gpurun_out/r2ac; mkdir -p $O
{
echo "== product"; timeout 200 python tools/stages.py large big
for w in 4 8 12; do echo "== sort begin bit $w"; timeout 200 python tools/stages.py large big 12=$w; done
} > $O/stages.log 2>&1
grep -v "^peak" $O/stages.log | cut -c1-200
