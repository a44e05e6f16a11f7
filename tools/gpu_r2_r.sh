#!/bin/bash
# round 2, GPU run R (N GPUs): multi-GPU parity, then the default bench line (dcp + the large block) on the final build
N=$1; O=gpurun_out/r2r; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 tools/dist_check.py > $O/dist_check_n$N.log 2>&1; echo "rc=$?" >> $O/dist_check_n$N.log
grep "dist_check\|rc=" $O/dist_check_n$N.log | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $N --steps 20 --warmup 5 > $O/dcp_n$N.json 2> $O/dcp_n$N.err; echo "bench rc=$?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for ln in open('gpurun_out/r2r/dcp_n%s.json'%N):
    if ln.startswith('{'): d=json.loads(ln)
print('dcp', d['ms_per_step'], '%.4g'%d['value'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d['large']['results'].items(): print(k, round(v['ms_per_step'],4), '%.4g'%v['value'], v.get('split_ms'))
PY
