#!/bin/bash
# Builds an experiment variant of the library: tools/build_variant.sh NAME [-DFLAG ...] -> build/variants/libswipe_b200_NAME.so
# (select it with SWB_LIBRARY=build/variants/libswipe_b200_NAME.so; build/ is git-ignored but travels with gpurun)
set -e
name=$1; shift
mkdir -p "$(dirname "$0")/../build/variants"
cd "$(dirname "$0")/../swipe_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o ../../build/variants/libswipe_b200_$name.so swb_api.cu swb_blastdb.cu swb_align.cu swb_scoring.cu swb_text.cu swb_ubench.cu
echo build/variants/libswipe_b200_$name.so
