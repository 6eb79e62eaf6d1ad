#!/bin/bash
V=build/variants/libswipe_b200_builder.so
SWB_LIBRARY=$V timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x --timeout 600 > gpurun_out/r2p_pytest.log 2>&1; echo "pytest(builder) rc=$?"; tail -3 gpurun_out/r2p_pytest.log
echo "== default"; timeout 600 python tools/tune_shapes.py 5000000 375 16x24 1 1 2>&1 | tail -1
echo "== builder warp"; SWB_LIBRARY=$V timeout 600 python tools/tune_shapes.py 5000000 375,1000 16x24,16x16,16x20 1 1 2>&1 | tail -6
echo "== default 1000"; timeout 600 python tools/tune_shapes.py 5000000 1000 16x16 1 1 2>&1 | tail -1
