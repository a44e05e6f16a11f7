#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; tail -4 $O/tests.log
echo "== small tail cache (default)" > $O/stages.log; timeout 200 python tools/stages.py demo dcp rpm fmr >> $O/stages.log 2>&1
echo "== 192 KB tail cache (12=0)" >> $O/stages.log; timeout 200 python tools/stages.py demo dcp rpm fmr 12=0 >> $O/stages.log 2>&1; cat $O/stages.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --large-block 0 > $O/dcp_n1.json 2> $O/dcp_n1.err
python - <<'PY'
import json
for ln in open('gpurun_out/r2q/dcp_n1.json'):
    if ln.startswith('{'): d=json.loads(ln)
print('dcp', d['ms_per_step'], d['run']['timed_regions_ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])
PY
