#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "offset_conv or bn_stats" 2>&1 | tail -3
SIGMA=2 timeout 120 python tools/trace_dcn.py > gpurun_out/r2_o_trace_s2.txt 2>&1
SIGMA=0.5 timeout 120 python tools/trace_dcn.py > gpurun_out/r2_o_trace_s05.txt 2>&1
cat gpurun_out/r2_o_trace_s2.txt gpurun_out/r2_o_trace_s05.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -c 1 -s 2 -o gpurun_out/r2_dcn_tc_v6a python tools/prof_conv.py fp16 dcn > gpurun_out/r2_ncu_dcn.log 2>&1
tail -2 gpurun_out/r2_ncu_dcn.log
