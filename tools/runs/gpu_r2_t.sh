#!/bin/bash
# 2-GPU check of the driver's scaling launch (forward arms + train record with the NCCL all-reduce inside the timed region)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err
head -c 300 gpurun_out/bench_r2_2gpu.json; echo; tail -3 gpurun_out/bench_r2_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_2gpu.json'))
print(d['value'], d['n_gpus'], d['train']['value'], d['train']['ms_per_step'], d['train']['config']['collective'], d['train']['config']['cuda_graph'], d['train']['config']['note'])
PY
