#!/bin/bash
# round 2, final 1-GPU run: the whole GPU suite, smoke, every config as a bench line (both arms)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2g_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2g_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2g_bench_reference.json 2>> gpurun_out/r2g_bench.err; echo "ref rc=$?"
for c in qlen100 qlen1000 qlen5000; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2g_bench_$c.json 2> gpurun_out/r2g_bench_$c.err; echo "$c rc=$?"; tail -2 gpurun_out/r2g_bench_$c.err
done
timeout 900 python bench.py --config qlen100 --batch 4 --steps 5 --warmup 3 > gpurun_out/r2g_bench_qlen100_batch4.json 2> gpurun_out/r2g_bench_qlen100_batch4.err; echo "batch rc=$?"
timeout 1500 python bench.py --config nt50m --steps 3 --warmup 3 > gpurun_out/r2g_bench_nt50m.json 2> gpurun_out/r2g_bench_nt50m.err; echo "nt50m rc=$?"; tail -2 gpurun_out/r2g_bench_nt50m.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench_under_ncu.log 2>&1
du -sh gpurun_out
