#!/bin/bash
O=gpurun_out/r2ay; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -2 $O/tests.log
timeout 120 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 300 python bench.py --workload demo --steps 20 --warmup 5 --no-cpu-baseline > $O/demo_n1.json 2> $O/demo_n1.err
timeout 300 python bench.py --workload demo --reuse-order 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/demo_reuse_n1.json 2> $O/demo_reuse_n1.err
python - <<'PY'
import json
for f in ('demo_n1','demo_reuse_n1'):
    for ln in open('gpurun_out/r2ay/%s.json'%f):
        if ln.startswith('{'): d=json.loads(ln)
    print(f, d['ms_per_step'], '%.4g'%d['value'])
PY
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py::test_golden_cases tests/test_gpu_parity.py::test_seeded_synthetic_vs_oracle -m gpu -q -x > $O/initcheck.log 2>&1; echo "initcheck rc=$?" >> $O/initcheck.log; tail -3 $O/initcheck.log
