#!/bin/bash
# TEST / DEBUG INFRASTRUCTURE ONLY: builds tests/simt_emu/_build/libnvb_emu.so (kernels under the CPU SIMT emulator).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/_build"
g++ -std=c++17 -O2 -fPIC -shared -ffp-contract=off -pthread -Wno-unused-variable -o "$HERE/_build/libnvb_emu.so" \
    "$HERE/cuda_emu.cpp" "$HERE/emu_kernels.cpp"
echo "built $HERE/_build/libnvb_emu.so"
