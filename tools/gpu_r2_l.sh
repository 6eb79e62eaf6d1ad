#!/bin/bash
# small shards (1/8 of the 5 M database): launch geometry
run() { echo "== $*"; env "$@" timeout 600 python tools/tune_shapes.py 625000 375 16x24 1 1 2>&1 | tail -1; }
run SWB_NOTHING=1
run SWB_OVERSUB=1
run SWB_OVERSUB=3
run SWB_OVERSUB=4
run SWB_CHUNK_BYTES=67108864
run SWB_CHUNK_BYTES=33554432
run SWB_CHUNK_BYTES=67108864 SWB_OVERSUB=1
