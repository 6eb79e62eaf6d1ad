#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "shape or edge or sink" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
timeout 600 python tools/tune_shapes.py 5000000 100 4x25,8x13 1 1 2>&1 | tail -2
timeout 600 python tools/tune_shapes.py 5000000 100 4x25 1 2 2>&1 | tail -1
timeout 600 python tools/tune_shapes.py 5000000 50,64 4x25 1,0 1 2>&1 | tail -4
