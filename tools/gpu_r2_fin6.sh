#!/bin/bash
# final commit: full GPU suite, then memcheck over the sanitizer subset + the session tests (the prep kernel's hand-off path)
O=gpurun_out/r2fin6; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; tail -2 $O/tests.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $(cat tools/sanitizer_subset.txt | tr '\n' ' ') tests/test_gpu_parity.py::test_static_target_session_keeps_the_targets_records -m gpu -q -x > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.log; tail -4 $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py::test_golden_cases tests/test_gpu_parity.py::test_large_cloud_path_vs_oracle tests/test_gpu_parity.py::test_super_node_level_with_degenerate_lines_and_without tests/test_gpu_parity.py::test_reused_order_gives_identical_results tests/test_gpu_shard.py::test_nccl_protocol_stages_emulated_on_one_gpu -m gpu -q -x > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log; tail -4 $O/sanitizer_racecheck.log
