#!/bin/bash
# round-2 evidence: ncu --set full captures of the deformable kernel and the four stage-4 conv classes (fp16 and tf32 arms),
# launch lists of one forward step per arm, library baselines
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -c 1 -s 2 -o gpurun_out/r2_dcn_tc_after python tools/prof_conv.py fp16 dcn > gpurun_out/r2_ncu_dcn_after.log 2>&1
for arm in fp16 tf32; do
  for cls in c48 c96 c192 c384; do
    timeout 300 ncu --set full --clock-control none -k regex:conv_ -c 1 -s 3 -o gpurun_out/r2_conv_${cls}_${arm} python tools/prof_conv.py $arm $cls > gpurun_out/r2_ncu_conv_${cls}_${arm}.log 2>&1
  done
done
for arm in tf32 fp16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_${arm}_step.csv python bench.py --profile-step --precision $arm > /dev/null 2>&1
  python tools/summarize_launches.py gpurun_out/r2_launches_${arm}_step.csv > gpurun_out/r2_launches_${arm}_step.txt 2>&1
done
timeout 300 python tools/bench_vs_libs.py gpurun_out/r2_vs_libs.json > gpurun_out/r2_vs_libs.txt 2>&1
ls -la gpurun_out | tail -30
