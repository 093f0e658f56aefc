#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -c 1 -s 2 -o gpurun_out/r2_dcn_tc_v2 python tools/prof_conv.py fp16 dcn > gpurun_out/r2_f_ncu.log 2>&1
tail -3 gpurun_out/r2_f_ncu.log
