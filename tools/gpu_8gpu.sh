#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${1:-8} --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus ${1:-8} --steps 5 --warmup 3 > gpurun_out/bench_${1:-8}gpu.json 2> gpurun_out/bench_${1:-8}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${1:-8}gpu.err
python - <<PY
import json
for ln in open('gpurun_out/bench_${1:-8}gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['clocks'])
PY
