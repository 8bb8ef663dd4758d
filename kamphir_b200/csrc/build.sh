#!/bin/bash
# Builds libphylok.so (host flattener + sm_100a kernels + C ABI) in-tree.  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${PK_OUT:-../libphylok.so}
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3,-ffp-contract=off,-Wall,-pthread"
$NVCC $FLAGS ${PK_DEFS:-} ${PK_PTXAS_V:+-Xptxas -v} -shared -o $OUT phylok_host.cpp phylok_dev.cu -lpthread
echo "built $(realpath $OUT)"
