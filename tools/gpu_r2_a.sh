#!/bin/bash
# round 2, call A: GPU tests + smoke + default bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench.err; cat gpurun_out/r2_bench.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err; echo "ref rc=$?"
cat gpurun_out/r2_bench_reference.json
nproc; lscpu | grep "Model name"
