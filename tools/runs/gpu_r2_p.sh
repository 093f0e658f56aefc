#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
BLOCKED=1 timeout 200 python tools/time_dcn.py 2>&1 | grep sigma
SIGMA=2 timeout 120 python tools/trace_dcn.py 2>&1 | tail -11
timeout 300 python tools/time_convs.py fp16 gpurun_out/r2_p_convs_fp16.json > gpurun_out/r2_p_convs_fp16.txt 2>&1
head -8 gpurun_out/r2_p_convs_fp16.txt | cut -c1-100
