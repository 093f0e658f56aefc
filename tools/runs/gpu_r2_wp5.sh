#!/bin/bash
# warp-private DCN kernel: DCN tests + timing + two ablations
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "dcn or deform or offset_conv" 2>&1 | tail -5 > gpurun_out/r2_wp5_pytest.txt
tail -3 gpurun_out/r2_wp5_pytest.txt
for a in 0 8 16; do
  echo "== FAMI_DCN_ABLATE=$a"
  FAMI_DCN_ABLATE=$a BLOCKED=1 timeout 100 python tools/time_dcn.py 2>&1 | grep sigma
done > gpurun_out/r2_wp5_time.txt 2>&1
cat gpurun_out/r2_wp5_time.txt
