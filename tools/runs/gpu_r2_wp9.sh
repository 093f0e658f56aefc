#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "dcn or deform or offset_conv" 2>&1 | tail -3
BLOCKED=1 timeout 100 python tools/time_dcn.py 2>&1 | grep sigma
