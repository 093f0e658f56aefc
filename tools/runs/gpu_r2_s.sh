#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/arm_errors.py > gpurun_out/r2_arm_errors.txt 2>&1
cat gpurun_out/r2_arm_errors.txt | tail -8
FAMI_STREAM_F32=1 timeout 600 python bench.py --precision fp16 --arms "" --no-train --no-reference-gpu --no-cpu-baseline --steps 10 > gpurun_out/r2_bench_fp16_stream.json 2> gpurun_out/r2_bench_fp16_stream.err
head -c 400 gpurun_out/r2_bench_fp16_stream.json; tail -2 gpurun_out/r2_bench_fp16_stream.err
