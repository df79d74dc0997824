#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  tests/_build/libintegration_b200.so = integration/B200Compressor.h (the maintainer-side binding of
# INTEGRATION.md section B) compiled against the reference's internal headers, linked with the unmodified reference library
# (oracle/_ref/libnvtt_ref.so, for its option classes) and with libnvtt_b200.so.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${NVTT_REFERENCE:-/root/reference}"
if [ ! -d "$REF/src/nvtt" ] || [ ! -f "$ROOT/oracle/_ref/libnvtt_ref.so" ]; then echo "reference not found (prebuilt tests/_build is used as is)"; exit 0; fi
mkdir -p "$HERE/_build/inc/nvtt"
cp "$ROOT/integration/B200Compressor.h" "$HERE/_build/inc/nvtt/B200Compressor.h"   # "dropped into src/nvtt": found as <nvtt/B200Compressor.h>
g++ -std=c++11 -O2 -w -fPIC -shared -I"$ROOT/oracle/_ref/gen" -I"$REF/src" -I"$REF/src/nvtt" -I"$REF/extern/poshlib" -I"$HERE/_build/inc" -I"$ROOT/include" \
    "$HERE/nvcompress/integration_harness.cpp" -o "$HERE/_build/libintegration_b200.so" \
    -L"$ROOT/oracle/_ref" -lnvtt_ref -L"$ROOT/nvidia-texture-tools_b200/lib" -lnvtt_b200 \
    -Wl,-rpath,"\$ORIGIN/../../oracle/_ref" -Wl,-rpath,"\$ORIGIN/../../nvidia-texture-tools_b200/lib"
echo "built $HERE/_build/libintegration_b200.so"
