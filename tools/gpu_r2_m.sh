#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_aux.py::test_demo_flow_with_the_reference_names > $O/tests.log 2>&1; tail -3 $O/tests.log
for cfg in "8000 12" "20000 12" "20000 25" "8000 25"; do echo "== lines/angle $cfg" >> $O/demo.log; timeout 200 python tools/demo_flaky.py 8 80 $cfg 2>&1 | grep "rot err" >> $O/demo.log; done; cat $O/demo.log
