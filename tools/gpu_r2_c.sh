#!/bin/bash
# round 2, call C: geometry-2 scan kernel: parity, then speed against geometry 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_geometry2.py -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_geom2.log 2>&1
echo "geom2 pytest rc=$?"; tail -12 gpurun_out/r2_pytest_geom2.log
# speed: 5 M subjects, 375 aa, both geometries (hybrid build)
timeout 600 python tools/tune_shapes.py 5000000 375 16x24 1 1 > gpurun_out/r2_tune_g1.jsonl 2>&1; cat gpurun_out/r2_tune_g1.jsonl
timeout 600 python tools/tune_shapes.py 5000000 375,1000,5000 16x20,16x21,16x24 1 2 > gpurun_out/r2_tune_g2.jsonl 2>&1; cat gpurun_out/r2_tune_g2.jsonl
timeout 600 python tools/tune_shapes.py 5000000 100 4x25 1 2 > gpurun_out/r2_tune_g2_100.jsonl 2>&1; cat gpurun_out/r2_tune_g2_100.jsonl
