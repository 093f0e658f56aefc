#!/bin/bash
# warp-private DCN kernel: one ncu --set full capture (sigma 2 launch)
mkdir -p gpurun_out
BLOCKED=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:dcn_wp_kernel -s 12 -c 1 \
  -o gpurun_out/r2_dcn_wp_v1 -f python tools/time_dcn.py > gpurun_out/r2_wp3_ncu.log 2>&1
tail -5 gpurun_out/r2_wp3_ncu.log
ls -la gpurun_out/
