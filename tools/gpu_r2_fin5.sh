#!/bin/bash
O=gpurun_out/r2fin5; mkdir -p $O
timeout 600 python bench.py > $O/dcp_default.json 2> $O/dcp_default.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in open('gpurun_out/r2fin5/dcp_default.json'):
    if ln.startswith('{'): d=json.loads(ln)
print(d['ms_per_step'], '%.4g'%d['value'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
print('dropin', d['dropin']['ms_per_step'], 'hooks', d['hooks']['ms_per_step'], 'ref gpu', d['reference_on_this_gpu']['ms_per_pair'], 'cpu', d['cpu_baseline']['value'])
for k,v in d['large']['results'].items(): print('   ', k, round(v['ms_per_step'],4))
PY
