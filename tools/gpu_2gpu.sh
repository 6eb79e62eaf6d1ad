#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -q -x --timeout 600 -k "two_gpus or protein_tsv" > gpurun_out/pytest_2gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
for ln in open('gpurun_out/bench_2gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'])
PY
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(1, d['value'], d['ms_per_step'], d['e2e'])"
