#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_geometry2.py -m gpu -q -x --timeout 600 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log
timeout 600 python tools/tune_shapes.py 5000000 1000,5000 16x16,32x20 1 1 2>&1 | tail -4
timeout 1500 python bench.py --config nt50m --steps 3 --warmup 3 > gpurun_out/r2j_bench_nt50m.json 2> gpurun_out/r2j_bench_nt50m.err; echo "nt50m rc=$?"; tail -2 gpurun_out/r2j_bench_nt50m.err
for c in qlen1000 qlen5000; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2j_bench_$c.json 2> gpurun_out/r2j_bench_$c.err; echo "$c rc=$?"; tail -2 gpurun_out/r2j_bench_$c.err
done
