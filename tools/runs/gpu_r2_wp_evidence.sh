#!/bin/bash
# evidence for the warp-private deformable kernel: timings, ablations, config-5 sweep, ncu capture, library baselines
mkdir -p gpurun_out
BLOCKED=1 timeout 200 python tools/time_dcn.py > gpurun_out/r2_wp_time_dcn.txt 2>&1
cat gpurun_out/r2_wp_time_dcn.txt
timeout 600 python tools/ablate_dcn_wp.py > gpurun_out/r2_dcn_wp_ablation.txt 2>&1
cat gpurun_out/r2_dcn_wp_ablation.txt
timeout 600 python tools/bench_dcn_sweep.py gpurun_out/r2_dcn_sweep.json > gpurun_out/r2_dcn_sweep.txt 2>&1
grep -c fp16_us gpurun_out/r2_dcn_sweep.txt
BLOCKED=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:dcn_wp_kernel -s 27 -c 1 \
  -o gpurun_out/r2_dcn_wp -f python tools/time_dcn.py > gpurun_out/r2_ncu_dcn_wp.log 2>&1
tail -2 gpurun_out/r2_ncu_dcn_wp.log
timeout 300 python tools/bench_vs_libs.py gpurun_out/r2_vs_libs.json > gpurun_out/r2_vs_libs.txt 2>&1
tail -3 gpurun_out/r2_vs_libs.txt
