#!/bin/bash
O=gpurun_out/r2at; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -2 $O/tests.log
timeout 300 python tools/stages.py demo dcp rpm fmr large big large8 large2 > $O/stages.log 2>&1; cut -c1-200 $O/stages.log
timeout 300 python bench.py --workload large --steps 20 --warmup 5 > $O/large_n1.json 2> $O/large_n1.err
python - <<'PY'
import json
for ln in open('gpurun_out/r2at/large_n1.json'):
    if ln.startswith('{'): d=json.loads(ln)
print('large', d['ms_per_step'], '%.4g'%d['value'], 'frac', d['roofline']['frac'])
for k,v in d['large']['results'].items(): print('   ', k, round(v['ms_per_step'],4), '%.4g'%v['value'])
PY
