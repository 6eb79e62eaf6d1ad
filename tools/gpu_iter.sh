#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
for v in "SWB_NO_SPEC=1" "SWB_X=1" "SWB_NO_SPEC=1" "SWB_X=1"; do
  echo "== $v"; env $v python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['gpu_launches'])"
done
