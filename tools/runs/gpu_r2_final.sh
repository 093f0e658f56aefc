#!/bin/bash
# round-2 final measurement pass: tests, smoke, bench (all arms + train + reference_gpu), reference arm, captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2_smoke.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_reference.json 2>> gpurun_out/bench_r2.err
tail -3 gpurun_out/r2_pytest_gpu.txt; tail -2 gpurun_out/r2_smoke.txt; head -c 300 gpurun_out/bench_r2.json; echo; tail -3 gpurun_out/bench_r2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -c 1 -s 2 -o gpurun_out/r2_dcn_tc_after python tools/prof_conv.py fp16 dcn > gpurun_out/r2_ncu_dcn_after.log 2>&1
for cls in c48 c96 c192 c384; do
  timeout 300 ncu --set full --clock-control none -k regex:conv_ -c 1 -s 3 -o gpurun_out/r2_conv_${cls}_tf32 python tools/prof_conv.py tf32 $cls > gpurun_out/r2_ncu_conv_${cls}_tf32.log 2>&1
done
for arm in tf32 fp16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_${arm}_step.csv python bench.py --profile-step --precision $arm > /dev/null 2>&1
  python tools/summarize_launches.py gpurun_out/r2_launches_${arm}_step.csv > gpurun_out/r2_launches_${arm}_step.txt 2>&1
done
timeout 300 python tools/bench_vs_libs.py gpurun_out/r2_vs_libs.json > gpurun_out/r2_vs_libs.txt 2>&1
timeout 300 python tools/time_convs.py tf32 gpurun_out/r2_convs_tf32.json > /dev/null 2>&1
timeout 300 python tools/time_convs.py fp16 gpurun_out/r2_convs_fp16.json > /dev/null 2>&1
ls gpurun_out | wc -l
