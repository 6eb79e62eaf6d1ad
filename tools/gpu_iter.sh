#!/bin/bash
# parity (quick subset or all) + shape tuning
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python tools/tune_shapes.py ${1:-2000000} ${2:-375} > gpurun_out/tune.jsonl 2> gpurun_out/tune.err
tail -3 gpurun_out/tune.err
cat gpurun_out/tune.jsonl
