#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 800 python tools/ablate_dcn.py > gpurun_out/r2_dcn_ablation.txt 2>&1; cat gpurun_out/r2_dcn_ablation.txt
SIGMA=2 timeout 100 python tools/trace_dcn.py 2>&1 | tail -11 > gpurun_out/r2_dcn_trace.txt; cat gpurun_out/r2_dcn_trace.txt
timeout 600 python tools/bench_dcn_sweep.py gpurun_out/r2_dcn_sweep.json > gpurun_out/r2_dcn_sweep.txt 2>&1; tail -3 gpurun_out/r2_dcn_sweep.txt | cut -c1-150
