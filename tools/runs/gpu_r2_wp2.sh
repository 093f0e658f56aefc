#!/bin/bash
# warp-private DCN kernel: timing ablations (results wrong by construction, only the time is read)
mkdir -p gpurun_out
for a in 0 1 4 8 16 128 20 148 156; do
  echo "== FAMI_DCN_ABLATE=$a"
  FAMI_DCN_ABLATE=$a BLOCKED=1 timeout 100 python tools/time_dcn.py 2>&1 | grep sigma
done > gpurun_out/r2_wp2_ablate.txt 2>&1
cat gpurun_out/r2_wp2_ablate.txt
