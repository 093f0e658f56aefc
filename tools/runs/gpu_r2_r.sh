#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "dcn or offset_conv" 2>&1 | tail -6
BLOCKED=1 timeout 200 python tools/time_dcn.py 2>&1 | grep sigma
timeout 600 python tools/bench_dcn_sweep.py gpurun_out/r2_dcn_sweep.json > gpurun_out/r2_dcn_sweep.txt 2>&1
tail -25 gpurun_out/r2_dcn_sweep.txt | cut -c1-220
