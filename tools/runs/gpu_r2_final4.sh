#!/bin/bash
# last pass of round 2: tests, smoke, bench line, reference arm, train-step call profile
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2_smoke.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_reference.json 2>> gpurun_out/bench_r2.err
timeout 600 python tools/profile_train.py > gpurun_out/r2_profile_train_calls.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu.txt; tail -2 gpurun_out/r2_smoke.txt; head -c 200 gpurun_out/bench_r2.json; echo; tail -2 gpurun_out/bench_r2.err; head -6 gpurun_out/r2_profile_train_calls.txt
