#!/bin/bash
# Builds the product library in-tree (C ABI + sm_100a kernels), twice from the same sources:
#   lib/libnvtt_b200.so           -fmad=false: never contract a*b+c into FMA (bit-exact parity with the reference's
#                                 non-FMA scalar code); -prec-div/-prec-sqrt stay at their IEEE defaults; no fast-math.
#   lib/libnvtt_b200_fastmath.so  -fmad=true -DNVB_FASTMATH: the opt-in variant OUTSIDE the parity contract (SURVEY 7.2
#                                 item 7): FMA contraction allowed, same C ABI (nvttb_build_variant() tells them apart).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/lib"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
COMMON="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -shared -cudart static"
$NVCC $COMMON -fmad=false ${NVB_PTXAS_V:+-Xptxas -v} -o "$HERE/lib/libnvtt_b200.so" "$HERE/csrc/capi.cu" &
strict=$!
if [ -z "$NVB_SKIP_FASTMATH" ]; then
    $NVCC $COMMON -fmad=true -DNVB_FASTMATH -o "$HERE/lib/libnvtt_b200_fastmath.so" "$HERE/csrc/capi.cu" &
    fast=$!
    wait $fast
fi
wait $strict
echo "built $HERE/lib/libnvtt_b200.so"
