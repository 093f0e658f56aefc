#!/bin/bash
# fast fp32-residual-stream epilogue: parity tests of the arms that use it, arm errors, step time
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -q -x 2>&1 | tail -3
timeout 600 python tools/arm_errors.py 2>&1 | tail -6
FAMI_STREAM_F32=1 timeout 600 python bench.py --precision fp16 --arms "" --no-train --no-reference-gpu --no-cpu-baseline --steps 10 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('fp16+stream', d['value'], d['ms_per_step'], d['launches_per_step'])"
