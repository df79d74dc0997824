#!/bin/bash
# Builds lib/libnvtt.so: the C++ nvtt:: API mirror, linked against the C ABI library only (no CUDA code here).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
LIB="$HERE/../lib"
mkdir -p "$LIB"
g++ -std=c++11 -O2 -fPIC -shared -fvisibility=hidden -ffp-contract=off -I"$HERE" \
    -o "$LIB/libnvtt.so" "$HERE/nvtt_host.cpp" -L"$LIB" -lnvtt_b200 -Wl,-rpath,'$ORIGIN'
echo "built $LIB/libnvtt.so"
