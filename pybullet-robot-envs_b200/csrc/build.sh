#!/bin/bash
# Build libb2env.so (the C-ABI library) for sm_100a, in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xptxas -v -shared -Xcompiler -fPIC -o libb2env.so b2env.cu "$@"
