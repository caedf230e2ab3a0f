#!/bin/sh
# Builds a variant of the pool kernel for A/B runs on the GPU box: scratch/libs/libblangpt_<name>.so
# (pool.cu recompiled with the given -D flags, linked with the engine object of the in-tree build).
# usage: tools/build_variant.sh <name> [nvcc flags...];   run with BLANGPT_LIB=scratch/libs/libblangpt_<name>.so
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; shift
CS="$ROOT/blangsdk_b200/csrc"
mkdir -p "$ROOT/scratch/libs" "$ROOT/scratch/obj"
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas --register-usage-level=8 "$@" \
  -c "$CS/pool.cu" -o "$ROOT/scratch/obj/pool_$NAME.o"
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -shared -gencode arch=compute_100a,code=sm_100a \
  -o "$ROOT/scratch/libs/libblangpt_$NAME.so" "$CS/build/engine.o" "$ROOT/scratch/obj/pool_$NAME.o"
echo "built scratch/libs/libblangpt_$NAME.so"
