#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2_bench_train.json 2> gpurun_out/r2_bench_train.err; head -c 200 gpurun_out/r2_bench_train.json; echo; tail -2 gpurun_out/r2_bench_train.err
