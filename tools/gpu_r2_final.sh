#!/bin/bash
# round 2, final verification on the committed state (1 GPU): full GPU suite, smoke, the driver's default bench line and the reference arm
O=gpurun_out/r2final; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 120 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py > $O/dcp_default.json 2> $O/dcp_default.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > $O/dcp_reference_arm.json 2> $O/dcp_reference_arm.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2final/dcp_default.json','gpurun_out/r2final/dcp_reference_arm.json'):
    d=None
    for ln in open(f):
        if ln.startswith('{'): d=json.loads(ln)
    print(f.split('/')[-1], d.get('ms_per_step'), '%.4g'%d.get('value',0), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'launches', d.get('gpu_launches'), 'steps', d.get('steps'), d.get('warmup'))
    print('  reference_on_this_gpu', d.get('reference_on_this_gpu'))
    print('  roofline frac', (d.get('roofline') or {}).get('frac'), 'clocks', d.get('clocks'))
PY
tail -5 $O/dcp_default.err
