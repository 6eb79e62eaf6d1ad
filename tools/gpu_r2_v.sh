#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 2 --config qlen1000 --steps 3 --warmup 3 > gpurun_out/r2v_bench_qlen1000_2gpu.json 2> gpurun_out/r2v.err; echo "qlen1000 N=2 rc=$?"; tail -3 gpurun_out/r2v.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 \
  bench.py --gpus 2 --config nt50m --nseq 5000000 --steps 3 --warmup 3 > gpurun_out/r2v_bench_nt5m_2gpu.json 2> gpurun_out/r2v2.err; echo "nt N=2 rc=$?"; tail -3 gpurun_out/r2v2.err | cut -c1-300
for f in gpurun_out/r2v_bench_*.json; do grep "^{" $f | cut -c1-200; done
