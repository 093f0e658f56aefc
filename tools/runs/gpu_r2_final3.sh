#!/bin/bash
# final pass of round 2 (warp-private deformable kernel, tf32 tensor-core wgrad): tests, smoke, bench line, reference arm,
# launch lists per arm, arm errors, wgrad timings
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2_smoke.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_reference.json 2>> gpurun_out/bench_r2.err
tail -3 gpurun_out/r2_pytest_gpu.txt; tail -2 gpurun_out/r2_smoke.txt; head -c 300 gpurun_out/bench_r2.json; echo; tail -3 gpurun_out/bench_r2.err
for arm in tf32 fp16s fp16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_${arm}_step.csv python bench.py --profile-step --precision $arm > /dev/null 2>&1
  python tools/summarize_launches.py gpurun_out/r2_launches_${arm}_step.csv > gpurun_out/r2_launches_${arm}_step.txt 2>&1
done
timeout 300 python tools/arm_errors.py > gpurun_out/r2_arm_errors.txt 2>&1
timeout 300 python tools/time_wgrad.py > gpurun_out/r2_wgrad_times.txt 2>&1
timeout 600 python tools/profile_train.py > gpurun_out/r2_profile_train_calls.txt 2>&1
tail -4 gpurun_out/r2_arm_errors.txt; head -12 gpurun_out/r2_profile_train_calls.txt
