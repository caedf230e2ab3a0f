#!/bin/sh
# Builds libblangpt.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU (about two minutes).
# Do not add --split-compile: it halves the build time but changes the code generated for the C2 kernel
# (measured: 2.78 -> 2.06 M evals/s).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
CCBIN="${CCBIN:-/usr/bin/g++}"
"$NVCC" -ccbin "$CCBIN" -std=c++17 -O3 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -shared $EXTRA_NVCC_FLAGS \
  -o "$HERE/../libblangpt.so" "$HERE/engine.cu"
