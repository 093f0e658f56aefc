#!/bin/bash
# warp-private DCN kernel v2: ablations + one ncu --set full capture (sigma 2 launch)
mkdir -p gpurun_out
for a in 0 1 8 16 24; do
  echo "== FAMI_DCN_ABLATE=$a"
  FAMI_DCN_ABLATE=$a BLOCKED=1 timeout 100 python tools/time_dcn.py 2>&1 | grep sigma
done > gpurun_out/r2_wp4_ablate.txt 2>&1
cat gpurun_out/r2_wp4_ablate.txt
BLOCKED=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:dcn_wp_kernel -s 27 -c 1 \
  -o gpurun_out/r2_dcn_wp_v2 -f python tools/time_dcn.py > gpurun_out/r2_wp4_ncu.log 2>&1
tail -3 gpurun_out/r2_wp4_ncu.log
