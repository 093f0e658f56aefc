#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_b_pytest_gpu.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_tf32_step.csv python bench.py --precision tf32 --profile-step > /dev/null 2>&1
timeout 300 python tools/time_convs.py tf32 gpurun_out/r2_b_convs_tf32.json > gpurun_out/r2_b_convs_tf32.txt 2>&1
timeout 300 python tools/time_convs.py fp16 gpurun_out/r2_b_convs_fp16.json > gpurun_out/r2_b_convs_fp16.txt 2>&1
tail -4 gpurun_out/r2_b_pytest_gpu.txt; python tools/summarize_launches.py gpurun_out/r2_launches_tf32_step.csv | head -20; cat gpurun_out/r2_b_convs_tf32.txt
