#!/bin/bash
O=gpurun_out/r2ax; mkdir -p $O
{
echo "== hand-off for B <= 4"; timeout 300 python tools/stages.py demo dcp
echo "== off (15=1)"; timeout 300 python tools/stages.py demo 15=1
} > $O/stages.log 2>&1; grep -v "^peak" $O/stages.log | cut -c1-100
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -2 $O/tests.log
for w in demo; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/${w}_n1.json 2> $O/${w}_n1.err; done
timeout 300 python bench.py --workload demo --reuse-order 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/demo_reuse_n1.json 2> $O/demo_reuse_n1.err
python - <<'PY'
import json
for f in ('demo_n1','demo_reuse_n1'):
    for ln in open('gpurun_out/r2ax/%s.json'%f):
        if ln.startswith('{'): d=json.loads(ln)
    print(f, d['ms_per_step'], '%.4g'%d['value'])
PY
