#!/bin/bash
# round 2, GPU run H (1 GPU): launch-geometry sweep of the super-node dense kernel at the line counts of an 8-rank / 2-rank shard
O=gpurun_out/r2h; mkdir -p $O
for w in 16 8 4 32; do for m in 32 16 64 128; do
  echo "== waves(9)=$w minnodes(3)=$m" >> $O/sweep.log
  timeout 120 python tools/stages.py large8 large2 9=$w 3=$m 2>&1 | grep -o '^[a-z0-9]* \|"dense": [0-9.]*\|"total": [0-9.]*' | paste -sd' ' >> $O/sweep.log
done; done
echo "== lpt=1 defaults" >> $O/sweep.log
timeout 120 python tools/stages.py large8 large2 6=1 2>&1 | grep -o '^[a-z0-9]* \|"dense": [0-9.]*\|"total": [0-9.]*' | paste -sd' ' >> $O/sweep.log
cat $O/sweep.log
