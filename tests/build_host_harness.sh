#!/bin/bash
# Test artefact: oracle/ref_harness.cpp (written against the REFERENCE's nvtt.h) compiled unchanged against the drop-in
# header and linked with our libnvtt.so -> tests/_build/libnvtt_b200_harness.so.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
PKG="$ROOT/nvidia-texture-tools_b200"
mkdir -p "$HERE/_build"
g++ -std=c++11 -O2 -fPIC -shared -I"$PKG/host" -o "$HERE/_build/libnvtt_b200_harness.so" \
    "$ROOT/oracle/ref_harness.cpp" -L"$PKG/lib" -lnvtt -lnvtt_b200 -Wl,-rpath,"$PKG/lib"
echo "built $HERE/_build/libnvtt_b200_harness.so"
