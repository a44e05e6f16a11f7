#!/bin/bash
O=gpurun_out/r2an; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "extreme_scales" > $O/tests_extreme.log 2>&1; tail -15 $O/tests_extreme.log
