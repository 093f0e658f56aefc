#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "dcn or deform or offset_conv" 2>&1 | tail -3
for m in 0 1; do
  echo "== FAMI_DCN_WP_PF=$m"
  FAMI_DCN_WP_PF=$m BLOCKED=1 timeout 100 python tools/time_dcn.py 2>&1 | grep sigma
done > gpurun_out/r2_wp8_time.txt 2>&1
echo "== FAMI_DCN_ABLATE=8" >> gpurun_out/r2_wp8_time.txt
FAMI_DCN_ABLATE=8 BLOCKED=1 timeout 100 python tools/time_dcn.py 2>&1 | grep sigma >> gpurun_out/r2_wp8_time.txt
cat gpurun_out/r2_wp8_time.txt
