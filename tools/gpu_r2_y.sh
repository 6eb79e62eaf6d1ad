#!/bin/bash
run() { echo "== $*"; env "$@" timeout 600 python tools/tune_shapes.py 5000000 375 16x24 1 1 2>&1 | tail -1 | cut -c60-140; }
run SWB_NOTHING=1
for cb in 402653184 536870912 671088640 1073741824; do
  run SWB_CHUNK_BYTES=$cb SWB_OVERSUB=1
  run SWB_CHUNK_BYTES=$cb SWB_OVERSUB=2
done
run SWB_NOTHING=1
