#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_model.py -q -x -k "dropin or calling or full_size or convention" -s 2>&1 | tail -25 > gpurun_out/r2_c_tests.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_c_bench.json 2> gpurun_out/r2_c_bench.err
FAMI_HALO_NORES=1 timeout 300 python tools/time_convs.py tf32 gpurun_out/r2_c_convs_tf32_nores.json > gpurun_out/r2_c_convs_tf32_nores.txt 2>&1
tail -8 gpurun_out/r2_c_tests.txt; tail -5 gpurun_out/r2_c_bench.err; head -c 3000 gpurun_out/r2_c_bench.json; head -6 gpurun_out/r2_c_convs_tf32_nores.txt
