#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
timeout 300 python tools/split_streams.py dcp > $O/split.log 2>&1; cat $O/split.log
timeout 300 python tools/split_streams.py fmr >> $O/split_fmr.log 2>&1; cat $O/split_fmr.log
