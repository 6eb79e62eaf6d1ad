#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tune_shapes.py 5000000 375 16x24 1 1 > gpurun_out/r2f_g1_375.jsonl 2>&1; cat gpurun_out/r2f_g1_375.jsonl
timeout 900 python tools/e2e_probe.py 1000 > gpurun_out/r2f_e2e_1000.jsonl 2>&1; cat gpurun_out/r2f_e2e_1000.jsonl
timeout 900 python tools/e2e_probe.py 375 > gpurun_out/r2f_e2e_375.jsonl 2>&1; cat gpurun_out/r2f_e2e_375.jsonl
