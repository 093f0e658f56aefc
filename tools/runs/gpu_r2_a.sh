#!/bin/bash
# round-2 GPU check A: TMA tf32 probe, tf32 conv + model parity, quick tf32 forward timing
mkdir -p gpurun_out
timeout 120 python tools/probe_tma_tf32.py > gpurun_out/r2_tma_tf32_probe.json 2> gpurun_out/r2_tma_tf32_probe.err
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "tf32" -s 2>&1 | tail -60 > gpurun_out/r2_a_tf32_ops.txt
timeout 600 python -m pytest tests/test_gpu_model.py -q -x -k "tf32 or eval_vs_reference" -s 2>&1 | tail -30 > gpurun_out/r2_a_tf32_model.txt
timeout 600 python bench.py --precision tf32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_a_bench_tf32.json 2> gpurun_out/r2_a_bench_tf32.err
tail -3 gpurun_out/r2_a_tf32_ops.txt gpurun_out/r2_a_tf32_model.txt; cat gpurun_out/r2_tma_tf32_probe.json; tail -c 1500 gpurun_out/r2_a_bench_tf32.json; tail -5 gpurun_out/r2_a_bench_tf32.err
