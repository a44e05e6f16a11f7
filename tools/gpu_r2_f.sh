#!/bin/bash
# round 2, GPU run F (1 GPU): target reuse + the fixed premise test, then the large workload
O=gpurun_out/r2f; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py tests/test_gpu_hooks.py -m gpu -q -x -k "static_target or more_d_entries or reused or twist_loss or fused or world_size" > $O/tests.log 2>&1; tail -5 $O/tests.log
timeout 300 python bench.py --workload large --steps 20 --warmup 5 --no-cpu-baseline > $O/large_n1.json 2> $O/large_n1.err; tail -c 1500 $O/large_n1.json; tail -3 $O/large_n1.err
