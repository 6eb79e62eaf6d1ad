#!/bin/bash
# last call of the round: what the driver runs (GPU suite, smoke, both bench arms) on the final tree
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2g_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench.err; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2>> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2g_bench.err
cat gpurun_out/r2g_bench.json | cut -c1-600
