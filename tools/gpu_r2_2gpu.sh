#!/bin/bash
# round 2: the strong-scaling bench (BASELINE configs[4]) on 2 GPUs, both arms, launched as the driver does
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench N=$N rc=$?"; tail -5 gpurun_out/r2_bench_${N}gpu.err; cat gpurun_out/r2_bench_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_${N}gpu.json 2>> gpurun_out/r2_bench_${N}gpu.err
echo "reference N=$N rc=$?"; cat gpurun_out/r2_bench_reference_${N}gpu.json
timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -q -x --timeout 600 -k "two_gpus or protein_tsv or nt_tsv or tblastn_tsv" > gpurun_out/r2_pytest_cli_${N}gpu.log 2>&1; echo "cli rc=$?"; tail -3 gpurun_out/r2_pytest_cli_${N}gpu.log
