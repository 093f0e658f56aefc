#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -x -k "dcn or tf32 or offset_conv" 2>&1 | tail -8 > gpurun_out/r2_g_tests.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_tf32_step.csv python bench.py --precision tf32 --profile-step > /dev/null 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --arms "" --no-train --no-reference-gpu --no-cpu-baseline > gpurun_out/r2_g_bench.json 2> gpurun_out/r2_g_bench.err
tail -4 gpurun_out/r2_g_tests.txt; python tools/summarize_launches.py gpurun_out/r2_launches_tf32_step.csv | head -12; head -c 300 gpurun_out/r2_g_bench.json; tail -3 gpurun_out/r2_g_bench.err
