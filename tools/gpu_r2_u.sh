#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "cast_nhwc or dcn_tf32 or dcn_tc" 2>&1 | tail -3
timeout 600 python bench.py --no-train --no-reference-gpu --no-cpu-baseline --steps 10 > gpurun_out/bench_tmp.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_tmp.json'))
for k,v in d['arms'].items(): print(k, round(v['value'],1), 'dcn', round(v['roofline']['us_per_launch'],1), round(v['roofline']['frac'],3))
PY
