#!/bin/bash
# the bench lines that read the final ncu summaries (roofline.executed): default dcp line and the large pair
O=gpurun_out/r2fin3; mkdir -p $O
timeout 600 python bench.py > $O/dcp_default.json 2> $O/dcp_default.err; echo "bench rc=$?"
timeout 300 python bench.py --workload large --steps 20 --warmup 5 > $O/large_n1.json 2> $O/large_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/dcp_n1.json 2> $O/dcp_n1.err
python - <<'PY'
import json
for f in ('dcp_default','dcp_n1','large_n1'):
    for ln in open('gpurun_out/r2fin3/%s.json'%f):
        if ln.startswith('{'): d=json.loads(ln)
    print(f, d['ms_per_step'], '%.4g'%d['value'], 'frac', d['roofline']['frac'], 'e2e', (d.get('e2e') or {}).get('ms_per_step'), (d.get('reference_on_this_gpu') or {}).get('ms_per_pair'))
    if d.get('large'):
        for k,v in d['large']['results'].items(): print('   ', k, round(v['ms_per_step'],4))
PY
