#!/usr/bin/env bash
# Tuning aid: builds build/variants/libb200rt_<name>.so with extra -D knobs (CUDA core and the wide-tree builder, whose
# node encoding is a compile-time choice; the other host objects are reused), to be A/B-tested on the GPU box with
# B200RT_LIB=build/variants/libb200rt_<name>.so python bench.py ...
#   tools/build_variant.sh l2b6 -DPPM_SCAN_LOADS=2 -DPPM_MIN_BLOCKS=6
set -euo pipefail
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off \
  "$@" -c ceng795_b200/csrc/b200rt_api.cu -o build/variants/api_$name.o
g++ -std=c++17 -O2 -fPIC -ffp-contract=off "$@" -c ceng795_b200/csrc/wide_bvh.cpp -o build/variants/wide_bvh_$name.o
others=$(ls build/*.cpp.o | grep -v wide_bvh.cpp.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/libb200rt_$name.so build/variants/api_$name.o \
  build/variants/wide_bvh_$name.o $others
echo build/variants/libb200rt_$name.so
