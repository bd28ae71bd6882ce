#!/usr/bin/env bash
# Tuning aid: builds build/variants/libb200rt_<name>.so with extra -D knobs for the CUDA core (host objects reused), to be
# A/B-tested on the GPU box with B200RT_LIB=build/variants/libb200rt_<name>.so python bench.py ...
#   tools/build_variant.sh l2b6 -DPPM_SCAN_LOADS=2 -DPPM_MIN_BLOCKS=6
set -euo pipefail
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off \
  "$@" -c ceng795_b200/csrc/b200rt_api.cu -o build/variants/api_$name.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/libb200rt_$name.so build/variants/api_$name.o \
  build/*.cpp.o
echo build/variants/libb200rt_$name.so
