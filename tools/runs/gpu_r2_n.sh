#!/bin/bash
# DCN (pixel, group)-lane kernel: tests + timing at three offset magnitudes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_n_pytest.txt
tail -5 gpurun_out/r2_n_pytest.txt
BLOCKED=1 timeout 200 python tools/time_dcn.py > gpurun_out/r2_n_time_dcn.txt 2>&1
cat gpurun_out/r2_n_time_dcn.txt
