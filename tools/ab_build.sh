#!/bin/bash
# build variant libraries: tools/ab_build.sh name "-DK4_MINBLOCKS=5" ...  -> build_variants/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o build_variants/lib_${name}.so taichi_three_b200/csrc/tina_b200.cu
