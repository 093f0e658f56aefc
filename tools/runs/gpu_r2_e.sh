#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "dcn" 2>&1 | tail -8 > gpurun_out/r2_e_tests.txt
BLOCKED=1 timeout 300 python tools/time_dcn.py > gpurun_out/r2_e_time_dcn.txt 2>&1
tail -4 gpurun_out/r2_e_tests.txt; cat gpurun_out/r2_e_time_dcn.txt
