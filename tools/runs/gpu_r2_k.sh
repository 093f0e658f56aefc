#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "conv_tc" > gpurun_out/r2_k_tests.txt 2>&1
tail -5 gpurun_out/r2_k_tests.txt
timeout 200 python tools/time_convs.py fp16 gpurun_out/r2_k_convs_fp16.json > gpurun_out/r2_k_convs_fp16.txt 2>&1
timeout 200 python tools/time_convs.py tf32 gpurun_out/r2_k_convs_tf32.json > gpurun_out/r2_k_convs_tf32.txt 2>&1
FAMI_TC_PAIR=0 timeout 200 python tools/time_convs.py fp16 gpurun_out/r2_k_convs_fp16_nopair.json > gpurun_out/r2_k_convs_fp16_nopair.txt 2>&1
tail -3 gpurun_out/r2_k_convs_fp16.txt
