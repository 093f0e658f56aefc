#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -x -k "fp16_stream or half_vs" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_fp16s_step.csv python bench.py --profile-step --precision fp16s > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_fp16s_step.csv | head -12
python - <<'PY'
import csv,re,collections
lines=[l for l in open('gpurun_out/r2_launches_fp16s_step.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
def us(r):
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
    return v/1e3 if u in ("ns","nsecond") else v*1e3 if u in ("ms","msecond") else v
# top 25 individual launches
top=sorted(rows,key=lambda r:-us(r))[:14]
for r in top: print("%8.1f us  grid %s block %s  %s"%(us(r), r.get("Grid Size"), r.get("Block Size"), re.sub(r"\(.*","",r["Kernel Name"])[-60:]))
PY
