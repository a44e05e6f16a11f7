#!/bin/bash
# round 2, GPU run P1 (1 GPU): ncu captures of the final kernels, launch lists, sanitizer logs
O=gpurun_out/r2p1; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_dcp python tools/prof_one.py dcp 3 > $O/ncu_dcp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_large python tools/prof_one.py large 3 > $O/ncu_large.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tail_kernel -s 2 -c 1 -o $O/prof_tail_dcp python tools/prof_one.py dcp 3 > $O/ncu_tail.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_bench_dcp.csv python bench.py --steps 2 --warmup 1 --graph 0 --no-cpu-baseline --large-block 0 --repeats 1 --e2e-repeats 1 > $O/bench_under_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_large.csv python tools/prof_one.py large 2 > $O/large_under_ncu.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $(cat tools/sanitizer_subset.txt | tr '\n' ' ') -m gpu -q -x > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.log; tail -4 $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py::test_golden_cases tests/test_gpu_parity.py::test_large_cloud_path_vs_oracle tests/test_gpu_shard.py::test_nccl_protocol_stages_emulated_on_one_gpu -m gpu -q -x > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log; tail -4 $O/sanitizer_racecheck.log
ls -la $O
