#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x --timeout 600 > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest.log
echo "== cp.async feed"; timeout 600 python tools/tune_shapes.py 5000000 1000 16x16,32x16,16x24,16x20 1 1 2>&1 | tail -4
timeout 600 python tools/tune_shapes.py 5000000 5000 32x20,16x24,16x20 1 1 2>&1 | tail -3
timeout 600 python tools/tune_shapes.py 5000000 375 16x24 1 1 2>&1 | tail -1
