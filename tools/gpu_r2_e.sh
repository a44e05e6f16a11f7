#!/bin/bash
# round 2, GPU run E (1 GPU): full suite on the final build, smoke, bench of every workload, final ncu captures, stage times
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $O/tests.log 2>&1; tail -12 $O/tests.log
timeout 120 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/dcp_n1.json 2> $O/dcp_n1.err; tail -c 400 $O/dcp_n1.json; echo
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/dcp_reference_arm.json 2> $O/dcp_reference_arm.err
timeout 300 python bench.py --workload large --steps 20 --warmup 5 > $O/large_n1.json 2> $O/large_n1.err
for w in rpm fmr demo; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/${w}_n1.json 2> $O/${w}_n1.err; done
timeout 300 python bench.py --workload demo --reuse-order 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/demo_reuse_n1.json 2> $O/demo_reuse_n1.err
timeout 300 python tools/stages.py demo dcp rpm fmr large > $O/stages.log 2>&1; cat $O/stages.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_dcp python tools/prof_one.py dcp 3 > $O/ncu_dcp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_large python tools/prof_one.py large 3 > $O/ncu_large.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tail_kernel -s 2 -c 1 -o $O/prof_tail_dcp python tools/prof_one.py dcp 3 > $O/ncu_tail.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_bench_dcp.csv python bench.py --steps 2 --warmup 1 --graph 0 --no-cpu-baseline --large-block 0 --repeats 1 --e2e-repeats 1 > $O/bench_under_ncu.log 2>&1
