#!/bin/bash
# Builds libjxl_b200.so (CUDA kernels for sm_100a + C ABI) in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  --shared -Xcompiler -fPIC,-Wall,-Wno-unused-parameter -Xptxas -v \
  --fmad=false -cudart static \
  -o libjxl_b200.so csrc/jxl_b200.cu 2> build.log || { cat build.log; exit 1; }
grep -E "Compiling entry|registers|spill" build.log | paste - - - | sed 's/ptxas info    : //g' | head -20 || true
