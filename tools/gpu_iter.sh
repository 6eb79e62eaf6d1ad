#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_blastdb.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
for v in "SWB_ONE_STREAM=1" "SWB_X=1" "SWB_ONE_STREAM=1" "SWB_X=1"; do
  echo "== $v"; env $v python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_ms'])"
done
