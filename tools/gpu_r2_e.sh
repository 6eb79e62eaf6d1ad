#!/bin/bash
# round 2, call E: batched queries (parity + speed), refreshed bench lines after the kernel changes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_batch.py tests/test_gpu_geometry2.py -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_e.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest_e.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.err; cat gpurun_out/r2e_bench.json
for b in 1 4; do
  timeout 900 python bench.py --config qlen100 --batch $b --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_qlen100_b$b.json 2> gpurun_out/r2e_bench_qlen100_b$b.err; echo "qlen100 batch $b rc=$?"
  tail -3 gpurun_out/r2e_bench_qlen100_b$b.err; cat gpurun_out/r2e_bench_qlen100_b$b.json
done
for c in qlen1000 qlen5000; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_$c.json 2> gpurun_out/r2e_bench_$c.err; echo "$c rc=$?"
  tail -3 gpurun_out/r2e_bench_$c.err; cat gpurun_out/r2e_bench_$c.json
done
