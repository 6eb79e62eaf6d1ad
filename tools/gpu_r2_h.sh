#!/bin/bash
# round 2, call H: after the slot-based multi-pass scratch and the multi-pass end-cell kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2h_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest_gpu.log
for c in qlen1000 qlen5000; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2h_bench_$c.json 2> gpurun_out/r2h_bench_$c.err; echo "$c rc=$?"; tail -2 gpurun_out/r2h_bench_$c.err
done
timeout 1500 python bench.py --config nt50m --steps 3 --warmup 3 > gpurun_out/r2h_bench_nt50m.json 2> gpurun_out/r2h_bench_nt50m.err; echo "nt50m rc=$?"; tail -2 gpurun_out/r2h_bench_nt50m.err
nvidia-smi --query-gpu=memory.used --format=csv
