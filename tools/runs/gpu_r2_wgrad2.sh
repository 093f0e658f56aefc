#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bwd_dense.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python tools/time_wgrad.py > gpurun_out/r2_wgrad_times.txt 2>&1
cat gpurun_out/r2_wgrad_times.txt
