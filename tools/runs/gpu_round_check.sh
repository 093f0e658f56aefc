set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r1_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r1_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_r1_fp16.json 2> gpurun_out/bench_err.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_err.txt
timeout 600 python bench.py --mode train --steps 5 > gpurun_out/bench_r1_train.json 2>> gpurun_out/bench_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1_launches_fp16_step.csv python bench.py --profile-step > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -c 1 -s 2 -o gpurun_out/r1_dcn_tc python tools/prof_conv.py fp16 dcn > /dev/null 2>&1
tail -2 gpurun_out/r1_pytest_gpu.txt; tail -1 gpurun_out/r1_smoke.txt; cat gpurun_out/bench_r1_fp16.json | cut -c1-400; cat gpurun_out/bench_r1_train.json | cut -c1-200; tail -3 gpurun_out/bench_err.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -c 1 -s 3 -o gpurun_out/r1_conv_halo_c48 python tools/prof_conv.py fp16 c48 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 -s 3 -o gpurun_out/r1_conv_tc_c192 python tools/prof_conv.py fp16 c192 > /dev/null 2>&1
timeout 300 python tools/bench_vs_libs.py gpurun_out/r1_vs_libs.json > /dev/null 2>&1
