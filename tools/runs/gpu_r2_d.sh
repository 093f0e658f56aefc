#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -x -k "tf32" 2>&1 | tail -8 > gpurun_out/r2_d_tests.txt
timeout 300 python tools/time_convs.py tf32 gpurun_out/r2_d_convs_tf32.json > gpurun_out/r2_d_convs_tf32.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --arms "" --no-train --no-reference-gpu --no-cpu-baseline > gpurun_out/r2_d_bench.json 2> gpurun_out/r2_d_bench.err
tail -4 gpurun_out/r2_d_tests.txt; cat gpurun_out/r2_d_convs_tf32.txt; head -c 400 gpurun_out/r2_d_bench.json
