#!/bin/bash
O=gpurun_out/r2san3; mkdir -p $O
T="tests/test_gpu_parity.py::test_reused_order_gives_identical_results tests/test_gpu_parity.py::test_golden_cases tests/test_gpu_aux.py::test_demo_flow_with_the_reference_names"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $T -m gpu -q -x > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log; tail -3 $O/memcheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py::test_reused_order_gives_identical_results tests/test_gpu_parity.py::test_golden_cases -m gpu -q -x > $O/initcheck.log 2>&1; echo "initcheck rc=$?" >> $O/initcheck.log; tail -3 $O/initcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py::test_reused_order_gives_identical_results -m gpu -q -x > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log; tail -3 $O/racecheck.log
