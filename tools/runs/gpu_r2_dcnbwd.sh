#!/bin/bash
# tf32 product phase of the deformable weight gradient: tests, per-call profile of the training step, train record
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train.py -q -m gpu -x -k "dcn_bwd or train" 2>&1 | tail -3
timeout 600 python tools/profile_train.py > gpurun_out/r2_profile_train_calls.txt 2>&1
head -8 gpurun_out/r2_profile_train_calls.txt
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2_bench_train.json 2> gpurun_out/r2_bench_train.err; head -c 250 gpurun_out/r2_bench_train.json; echo; tail -2 gpurun_out/r2_bench_train.err
