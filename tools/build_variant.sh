#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFOO=1 ...]   -> tools/variants/NAME.so (+ NAME.log with ptxas -v)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared --fmad=true -Xptxas -v "$@" \
  -o tools/variants/$name.so crystalenergygrids.jl_b200/csrc/ceg_cuda.cu > tools/variants/$name.log 2>&1
grep -A2 "grid_build_tiled_kernelILi256ELi1ELb1ELb1" tools/variants/$name.log | grep -E "spill|Used" 
