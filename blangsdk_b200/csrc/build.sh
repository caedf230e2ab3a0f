#!/bin/sh
# Builds libblangpt.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
CCBIN="${CCBIN:-/usr/bin/g++}"
"$NVCC" -ccbin "$CCBIN" -std=c++17 -O3 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -shared $EXTRA_NVCC_FLAGS \
  -o "$HERE/../libblangpt.so" "$HERE/engine.cu"
