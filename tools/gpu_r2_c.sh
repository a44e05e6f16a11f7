#!/bin/bash
# round 2, GPU run C (1 GPU): A/B of the large-cloud variants, new sampler, sanitizer log, node_kernel profile
O=gpurun_out/r2c; mkdir -p $O
V=a-robust-registration-loss_b200/build/variants
timeout 600 python -m pytest tests/test_gpu_aux.py -m gpu -q -x -k "sampler or demo" > $O/tests_sampler.log 2>&1; tail -3 $O/tests_sampler.log
timeout 120 python tools/sampler_time.py > $O/sampler_time.log 2>&1; cat $O/sampler_time.log
for v in default node3 node4 pf2 pf12; do
  if [ $v = default ]; then unset RRL_LIB_PATH; else export RRL_LIB_PATH=$PWD/$V/librrl_b200_$v.so; fi
  echo "== $v" >> $O/stages_variants.log
  timeout 200 python tools/stages.py large big >> $O/stages_variants.log 2>&1
  echo "== $v lpt=1" >> $O/stages_variants.log
  timeout 200 python tools/stages.py large big 6=1 >> $O/stages_variants.log 2>&1
done
unset RRL_LIB_PATH
cat $O/stages_variants.log
export RRL_LIB_PATH=$PWD/$V/librrl_b200_pf12.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large or super or mid or reused" > $O/tests_pf12.log 2>&1; tail -3 $O/tests_pf12.log
unset RRL_LIB_PATH
timeout 300 ncu --set full --clock-control none --import-source on -k regex:node_kernel -s 2 -c 1 -o $O/prof_node_large python tools/prof_one.py large 3 > $O/ncu_node.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_shard.py tests/test_gpu_hooks.py -m gpu -q -x -k "fused_peer_exchange_tail and 3001 or stale or expmap or chamfer or allreduce" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.log; tail -8 $O/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_shard.py -m gpu -q -x -k "fused_peer_exchange_tail and 3001" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log; tail -8 $O/sanitizer_racecheck.log
