#!/bin/bash
# last call of the round: the GPU suite and the default bench line on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2g_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2g_bench.err
