#!/bin/bash
# N-GPU check of the driver's scaling launch (N = number of visible GPUs)
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
tail -2 gpurun_out/bench_r2_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2_${N}gpu.json'))
print(d['n_gpus'], d['value'], d['e2e']['value'], 'train', d['train']['value'], d['train']['ms_per_step'], d['train']['config']['collective'], d['train']['config']['cuda_graph'])
PY
