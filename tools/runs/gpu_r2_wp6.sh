#!/bin/bash
mkdir -p gpurun_out
for a in 0 32; do
  echo "== FAMI_DCN_ABLATE=$a"
  FAMI_DCN_ABLATE=$a BLOCKED=1 timeout 100 python tools/time_dcn.py 2>&1 | grep sigma
done > gpurun_out/r2_wp6_time.txt 2>&1
cat gpurun_out/r2_wp6_time.txt
