#!/bin/bash
for os in 1 2 3; do
echo "== SWB_OVERSUB=$os full"; SWB_OVERSUB=$os timeout 600 python tools/e2e_probe.py 375 5000000 2>&1 | head -1
echo "== SWB_OVERSUB=$os 1/8"; SWB_OVERSUB=$os timeout 600 python tools/e2e_probe.py 375 625000 2>&1 | head -1
done
