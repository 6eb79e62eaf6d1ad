#!/bin/bash
# blast-ingest GPU tests + tile microbenchmark
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blastdb.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 tools/ubench/tile > gpurun_out/ubench_tile.txt 2>&1; cat gpurun_out/ubench_tile.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['gpu_launches'])"
