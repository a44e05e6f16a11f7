#!/bin/bash
# round 2, GPU run P2 (1 GPU): smoke, the bench of every workload, stage times, queue counters on the final build
O=gpurun_out/r2p2; mkdir -p $O
timeout 120 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/dcp_n1.json 2> $O/dcp_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/dcp_reference_arm.json 2> $O/dcp_reference_arm.err
timeout 300 python bench.py --workload large --steps 20 --warmup 5 > $O/large_n1.json 2> $O/large_n1.err
for w in rpm fmr demo; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/${w}_n1.json 2> $O/${w}_n1.err; done
timeout 300 python bench.py --workload demo --reuse-order 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/demo_reuse_n1.json 2> $O/demo_reuse_n1.err
timeout 300 python bench.py --workload dcp --graph 0 --steps 20 --warmup 5 --no-cpu-baseline --large-block 0 > $O/dcp_n1_eager.json 2> $O/dcp_n1_eager.err
timeout 300 python tools/stages.py demo dcp rpm fmr large big large8 large2 > $O/stages.log 2>&1
timeout 100 python tools/sampler_time.py > $O/sampler_time.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p2/*.json')):
    d=None
    for ln in open(f):
        if ln.startswith('{'): d=json.loads(ln)
    if d is None: print(f,'NO LINE'); continue
    r=d.get('roofline') or {}
    print(f.split('/')[-1], d.get('ms_per_step'), '%.4g'%d.get('value',0), 'frac', r.get('frac'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
    if d.get('large'):
        for k,v in d['large']['results'].items(): print('   ', k, round(v['ms_per_step'],4), '%.4g'%v['value'])
PY
cat $O/stages.log; tail -3 $O/sampler_time.log
