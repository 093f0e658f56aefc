#!/bin/bash
# session-2 baseline profile: launch lists of one step per arm, per-class conv timings, full ncu capture of the DCN kernel
mkdir -p gpurun_out
for arm in tf32 fp16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_${arm}_step.csv python bench.py --profile-step --precision $arm > /dev/null 2>&1
  python tools/summarize_launches.py gpurun_out/r2_launches_${arm}_step.csv > gpurun_out/r2_launches_${arm}_step.txt 2>&1
done
timeout 300 python tools/time_convs.py tf32 gpurun_out/r2_convs_tf32.json > gpurun_out/r2_convs_tf32.txt 2>&1
timeout 300 python tools/time_convs.py fp16 gpurun_out/r2_convs_fp16.json > gpurun_out/r2_convs_fp16.txt 2>&1
BLOCKED=1 timeout 200 python tools/time_dcn.py > gpurun_out/r2_time_dcn.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -c 1 -s 2 -o gpurun_out/r2_dcn_tc_before python tools/prof_conv.py fp16 dcn > gpurun_out/r2_ncu_dcn.log 2>&1
head -12 gpurun_out/r2_launches_tf32_step.txt; head -8 gpurun_out/r2_launches_fp16_step.txt; cat gpurun_out/r2_time_dcn.txt
