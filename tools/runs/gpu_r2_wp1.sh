#!/bin/bash
# warp-private DCN kernel (dcn_wp.cu): DCN tests + timing, A/B against the tcgen05 kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "dcn or deform" 2>&1 | tail -15 > gpurun_out/r2_wp1_pytest.txt
tail -8 gpurun_out/r2_wp1_pytest.txt
BLOCKED=1 timeout 200 python tools/time_dcn.py > gpurun_out/r2_wp1_time_dcn.txt 2>&1
cat gpurun_out/r2_wp1_time_dcn.txt
FAMI_DCN_WP=0 BLOCKED=1 timeout 200 python tools/time_dcn.py > gpurun_out/r2_wp1_time_dcn_tc.txt 2>&1
cat gpurun_out/r2_wp1_time_dcn_tc.txt
