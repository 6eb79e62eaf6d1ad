#!/bin/bash
# parity (all GPU tests) + shape tuning: gpu_iter.sh [nseq] [qlens] [shapes] [modes]
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python tools/tune_shapes.py ${1:-2000000} ${2:-375} ${3:-16x24,32x12,16x12,8x16} ${4:-1,0} > gpurun_out/tune.jsonl 2> gpurun_out/tune.err
tail -3 gpurun_out/tune.err
cat gpurun_out/tune.jsonl
