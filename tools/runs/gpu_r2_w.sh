#!/bin/bash
# double-buffered wgrad: parity tests + per-call profile of the training step
timeout 600 python -m pytest tests/test_gpu_bwd_dense.py tests/test_gpu_train.py -q -x 2>&1 | tail -3
timeout 300 python tools/profile_train.py 32 tf32 tf32 2>&1 | grep -E "^step|wgrad" | head -12
FAMI_WGRAD_SINGLE=1 timeout 300 python tools/profile_train.py 32 tf32 tf32 2>&1 | grep -E "^step|wgrad" | head -4
