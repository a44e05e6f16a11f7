#!/bin/bash
# round 2, GPU run L (1 GPU): guards from the measured thr_max (1.25x over the derived bound), padded strides restored
O=gpurun_out/r2l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 200 python tools/stages.py demo dcp rpm fmr large big large8 > $O/stages.log 2>&1; cat $O/stages.log
RRL_LIB_PATH=$PWD/a-robust-registration-loss_b200/build/variants/librrl_b200_counters.so timeout 200 python tools/counters.py large dcp > $O/counters.log 2>&1; cat $O/counters.log
timeout 300 python bench.py --workload large --steps 20 --warmup 5 --no-cpu-baseline > $O/large_n1.json 2> $O/large_n1.err
python - <<'PY'
import json
for ln in open('gpurun_out/r2l/large_n1.json'):
    if ln.startswith('{'): d=json.loads(ln)
print({k:(v['ms_per_step']) for k,v in d['large']['results'].items()}, d['roofline']['kernel_ms'])
PY
