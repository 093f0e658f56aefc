#!/bin/bash
# tf32 tensor-core wgrad: tests, kernel timings, training-step record
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bwd_dense.py tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python tools/time_wgrad.py > gpurun_out/r2_wgrad_times.txt 2>&1
cat gpurun_out/r2_wgrad_times.txt
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2_bench_train.json 2> gpurun_out/r2_bench_train.err; head -c 600 gpurun_out/r2_bench_train.json; tail -3 gpurun_out/r2_bench_train.err
