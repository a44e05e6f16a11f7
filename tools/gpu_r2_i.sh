#!/bin/bash
# round 2, GPU run I (1 GPU): sampler with finer phases, demo convergence spread, the fixed premise test
O=gpurun_out/r2i; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_aux.py tests/test_gpu_parity.py -m gpu -q -x -k "sampler or more_d_entries or demo" > $O/tests.log 2>&1; tail -4 $O/tests.log
timeout 120 python tools/sampler_time.py > $O/sampler_time.log 2>&1; cat $O/sampler_time.log
timeout 300 python tools/demo_flaky.py 6 60 > $O/demo_flaky.log 2>&1; cat $O/demo_flaky.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --large-block 0 > $O/dcp_n1.json 2> $O/dcp_n1.err; python - <<'PY'
import json
for ln in open('gpurun_out/r2i/dcp_n1.json'):
    if ln.startswith('{'): d=json.loads(ln)
print(d['ms_per_step'], d['lines_sampled_on_device'])
PY
