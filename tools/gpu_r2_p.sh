#!/bin/bash
# round 2, GPU run P (1 GPU): final ncu captures, then the bench of every workload on the final build
O=gpurun_out/r2p; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_dcp python tools/prof_one.py dcp 3 > $O/ncu_dcp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_large python tools/prof_one.py large 3 > $O/ncu_large.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_bench_dcp.csv python bench.py --steps 2 --warmup 1 --graph 0 --no-cpu-baseline --large-block 0 --repeats 1 --e2e-repeats 1 > $O/bench_under_ncu.log 2>&1
timeout 120 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/dcp_n1.json 2> $O/dcp_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/dcp_reference_arm.json 2> $O/dcp_reference_arm.err
timeout 300 python bench.py --workload large --steps 20 --warmup 5 > $O/large_n1.json 2> $O/large_n1.err
for w in rpm fmr demo; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/${w}_n1.json 2> $O/${w}_n1.err; done
timeout 300 python bench.py --workload demo --reuse-order 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/demo_reuse_n1.json 2> $O/demo_reuse_n1.err
timeout 300 python bench.py --workload dcp --graph 0 --steps 20 --warmup 5 --no-cpu-baseline --large-block 0 > $O/dcp_n1_eager.json 2> $O/dcp_n1_eager.err
ls -la $O | head -30
