#!/bin/bash
# round 2, call D: parity after the row-loop rewrite, then speed of both geometries
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_geometry2.py tests/test_gpu_golden.py -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_d.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_d.log
timeout 900 python tools/tune_shapes.py 5000000 375 16x24,16x20 1 1 > gpurun_out/r2d_g1_375.jsonl 2>&1; cat gpurun_out/r2d_g1_375.jsonl
timeout 900 python tools/tune_shapes.py 5000000 100 8x13 1 1 > gpurun_out/r2d_g1_100.jsonl 2>&1; cat gpurun_out/r2d_g1_100.jsonl
timeout 900 python tools/tune_shapes.py 5000000 1000 32x16,32x32,16x24 1 1 > gpurun_out/r2d_g1_1000.jsonl 2>&1; cat gpurun_out/r2d_g1_1000.jsonl
timeout 900 python tools/tune_shapes.py 5000000 5000 32x20,32x16,32x24 1 1 > gpurun_out/r2d_g1_5000.jsonl 2>&1; cat gpurun_out/r2d_g1_5000.jsonl
for st in 1 0; do
  echo "stagger $st"
  SWB_STAGGER=$st timeout 900 python tools/tune_shapes.py 5000000 375,1000 16x24,16x21 1 2 > gpurun_out/r2d_g2_s$st.jsonl 2>&1; cat gpurun_out/r2d_g2_s$st.jsonl
done
SWB_STAGGER=1 timeout 900 python tools/tune_shapes.py 5000000 100 4x25 1 2 > gpurun_out/r2d_g2_100.jsonl 2>&1; cat gpurun_out/r2d_g2_100.jsonl
