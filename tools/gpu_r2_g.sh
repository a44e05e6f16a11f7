#!/bin/bash
# round 2, GPU run G (8 GPUs): multi-GPU parity incl. the peer exchange, then the bench at N=8 (dcp + the large block)
O=gpurun_out/r2g; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tools/dist_check.py > $O/dist_check_n8.log 2>&1; echo "rc=$?" >> $O/dist_check_n8.log
grep "dist_check\|rc=" $O/dist_check_n8.log | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 5 > $O/dcp_n8.json 2> $O/dcp_n8.err; echo "bench rc=$?"
tail -c 2500 $O/dcp_n8.json; tail -3 $O/dcp_n8.err
