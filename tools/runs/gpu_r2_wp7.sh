#!/bin/bash
mkdir -p gpurun_out
BLOCKED=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:dcn_wp_kernel -s 27 -c 1 \
  -o gpurun_out/r2_dcn_wp_v3 -f python tools/time_dcn.py > gpurun_out/r2_wp7_ncu.log 2>&1
tail -3 gpurun_out/r2_wp7_ncu.log
