#!/bin/bash
# round 2, GPU run B (2 GPUs): multi-GPU parity incl. the peer exchange over real NVLink, then the bench at N=2
O=gpurun_out/r2b; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > $O/dist_check_n2.log 2>&1; echo "rc=$?" >> $O/dist_check_n2.log
tail -5 $O/dist_check_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $O/dcp_n2.json 2> $O/dcp_n2.err; echo "bench rc=$?"
tail -c 3000 $O/dcp_n2.json; tail -5 $O/dcp_n2.err
