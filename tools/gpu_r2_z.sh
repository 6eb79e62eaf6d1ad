#!/bin/bash
timeout 600 python tools/tune_shapes.py 5000000 375 16x24 1 1 2>&1 | tail -1 | cut -c1-130
timeout 600 python tools/tune_shapes.py 5000000 1000 16x16 1 1 2>&1 | tail -1 | cut -c1-130
timeout 600 python tools/tune_shapes.py 5000000 5000 32x20 1 1 2>&1 | tail -1 | cut -c1-130
echo "== 256 MB chunks"
SWB_CHUNK_BYTES=268435456 timeout 600 python tools/tune_shapes.py 5000000 1000 16x16 1 1 2>&1 | tail -1 | cut -c1-130
SWB_CHUNK_BYTES=268435456 timeout 600 python tools/tune_shapes.py 5000000 5000 32x20 1 1 2>&1 | tail -1 | cut -c1-130
