#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -q -s > gpurun_out/r2_h_tests_full.txt 2>&1
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2_h_train.json 2> gpurun_out/r2_h_train.err
grep -n "graph vs eager\|resume:\|unfrozen train\|tf32-head\|grad dev\|^E  \|passed\|failed" gpurun_out/r2_h_tests_full.txt | head -60; head -c 900 gpurun_out/r2_h_train.json
