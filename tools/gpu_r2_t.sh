#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -k "not integration and not cli" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_pytest.log
bash tools/gpu_r2_s.sh 2>&1 | tail -12
timeout 600 python tools/e2e_probe.py 375 625000 2>&1 | head -1
timeout 600 python tools/e2e_probe.py 375 5000000 2>&1 | head -1
