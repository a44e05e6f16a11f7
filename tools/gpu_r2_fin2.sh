#!/bin/bash
# final ncu captures (dense kernels, node kernel) and launch list of the large pair on the final build
O=gpurun_out/r2fin2; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_large python tools/prof_one.py large 3 > $O/ncu_large.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 2 -c 1 -o $O/prof_dense_dcp python tools/prof_one.py dcp 3 > $O/ncu_dcp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:node_kernel -s 2 -c 1 -o $O/prof_node_large python tools/prof_one.py large 3 > $O/ncu_node.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_large.csv python tools/prof_one.py large 2 > $O/large_under_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_bench_dcp.csv python bench.py --steps 2 --warmup 1 --graph 0 --no-cpu-baseline --large-block 0 --repeats 1 --e2e-repeats 1 > $O/bench_under_ncu.log 2>&1
ls $O
