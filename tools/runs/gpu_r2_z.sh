#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py tests/test_gpu_train.py tests/test_gpu_dropin.py -q -x 2>&1 | tail -3
timeout 300 python tools/arm_errors.py 2>&1 | grep -E "^tf32|^fp32 "
timeout 600 python bench.py --no-train --no-reference-gpu --no-cpu-baseline --steps 10 --arms "" 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['ms_per_step'], d['launches_per_step'])"
