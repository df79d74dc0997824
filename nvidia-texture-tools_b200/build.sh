#!/bin/bash
# Builds the product library in-tree: nvidia-texture-tools_b200/lib/libnvtt_b200.so (C ABI + sm_100a kernels).
# -fmad=false: never contract a*b+c into FMA (bit-exact parity with the reference's non-FMA scalar code);
# -prec-div/-prec-sqrt stay at their IEEE defaults; no fast-math.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/lib"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
    -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -shared -cudart static \
    ${NVB_PTXAS_V:+-Xptxas -v} \
    -o "$HERE/lib/libnvtt_b200.so" "$HERE/csrc/capi.cu"
echo "built $HERE/lib/libnvtt_b200.so"
