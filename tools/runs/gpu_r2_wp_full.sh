#!/bin/bash
# full validation with the warp-private deformable kernel in the model: tests, smoke, bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_wp_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2_wp_smoke.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_wp.json 2> gpurun_out/bench_r2_wp.err
tail -3 gpurun_out/r2_wp_pytest_gpu.txt; tail -2 gpurun_out/r2_wp_smoke.txt; head -c 400 gpurun_out/bench_r2_wp.json; echo; tail -3 gpurun_out/bench_r2_wp.err
