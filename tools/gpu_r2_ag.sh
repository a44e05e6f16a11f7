#!/bin/bash
O=gpurun_out/r2ag; mkdir -p $O
timeout 300 python tools/stages.py large large8 large2 big > $O/stages.log 2>&1; cat $O/stages.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large or big or degenerate or brute" > $O/tests.log 2>&1; tail -3 $O/tests.log
