#!/bin/bash
mkdir -p gpurun_out
FAMI_HALO_PAIR=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "conv_tc" > gpurun_out/r2_l_tests.txt 2>&1
tail -5 gpurun_out/r2_l_tests.txt
FAMI_HALO_PAIR=1 timeout 200 python tools/time_convs.py fp16 gpurun_out/r2_l_convs_fp16_pair.json > gpurun_out/r2_l_convs_fp16_pair.txt 2>&1
FAMI_HALO_PAIR=1 timeout 200 python tools/time_convs.py tf32 gpurun_out/r2_l_convs_tf32_pair.json > gpurun_out/r2_l_convs_tf32_pair.txt 2>&1
tail -3 gpurun_out/r2_l_convs_tf32_pair.txt
