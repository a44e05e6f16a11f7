#!/bin/bash
# round 2, GPU run D (1 GPU): node_kernel occupancy A/B, level-2 preload A/B, sharded sampler test, sanitizer logs, full suite
O=gpurun_out/r2d; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
for v in default node5 node6 node8 pre pre2; do
  if [ $v = default ]; then unset RRL_LIB_PATH; else export RRL_LIB_PATH=$PWD/$V/librrl_b200_$v.so; fi
  echo "== $v" >> $O/stages_variants.log
  timeout 200 python tools/stages.py large big >> $O/stages_variants.log 2>&1
done
unset RRL_LIB_PATH
cat $O/stages_variants.log
for v in pre pre2; do RRL_LIB_PATH=$PWD/$V/librrl_b200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large or super or mid or reused" > $O/tests_$v.log 2>&1; tail -2 $O/tests_$v.log; done
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > $O/tests.log 2>&1; tail -12 $O/tests.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $(cat tools/sanitizer_subset.txt | tr '\n' ' ') -m gpu -q -x > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.log; tail -6 $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py::test_golden_cases tests/test_gpu_parity.py::test_large_cloud_path_vs_oracle tests/test_gpu_shard.py::test_nccl_protocol_stages_emulated_on_one_gpu -m gpu -q -x > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log; tail -6 $O/sanitizer_racecheck.log
