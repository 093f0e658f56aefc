#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
BN_BY_SHAPE=1 TOPN=70 timeout 600 python tools/profile_train.py 2>&1 | grep -E "^step|bn_apply_act N=160 96x72 C=48|bn_stats rows=1105920 C=48|bn_apply_act N=160 48x36 C=96" | head -8
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2_bench_train.json 2> gpurun_out/r2_bench_train.err; head -c 200 gpurun_out/r2_bench_train.json; echo; tail -2 gpurun_out/r2_bench_train.err
