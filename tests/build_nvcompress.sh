#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Builds the reference's nvcompress tool (src/nvtt/tools/compress.cpp, UNCHANGED, compiled where it
# lies under /root/reference) twice:
#   tests/_build/nvcompress_b200 - against the drop-in header host/nvtt/nvtt.h and lib/libnvtt.so + libnvtt_b200.so; the
#                                  reference's nvcore / nvimage / nvmath objects are linked in for FILE IO ONLY
#                                  (PNG / TGA / DDS reading, paths, the output stream) - none of its nvtt objects.
#   tests/_build/nvcompress_ref  - against the reference's own library objects (the pinned parity build, oracle/_ref).
# tests/test_gpu_nvcompress.py runs both on the same files and compares the .dds / .ktx output byte for byte.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${NVTT_REFERENCE:-/root/reference}"
OBJ="$ROOT/oracle/_ref/obj_pinned"
OUT="$HERE/_build"
if [ ! -d "$REF/src/nvtt/tools" ] || [ ! -d "$OBJ" ]; then echo "reference sources / objects not found (prebuilt tests/_build is used as is)"; exit 0; fi
mkdir -p "$OUT"
GEN="$ROOT/oracle/_ref/gen"
REFINC="-I$GEN -I$REF/src -I$REF/extern/poshlib"
IO_OBJS=$(ls $OBJ/src_nvcore_*.o $OBJ/src_nvimage_*.o $OBJ/src_nvmath_*.o $OBJ/src_bc6h_*.o $OBJ/src_bc7_*.o $OBJ/src_nvthread_*.o $OBJ/posh.o)
ALL_OBJS=$(ls $OBJ/*.o | grep -v ref_harness.o)
LIB="$ROOT/nvidia-texture-tools_b200/lib"
# ours: -I host first so that <nvtt/nvtt.h> is the drop-in header
g++ -std=c++11 -O2 -w -I"$ROOT/nvidia-texture-tools_b200/host" $REFINC \
    "$REF/src/nvtt/tools/compress.cpp" "$HERE/nvcompress/loader_shim.cpp" $IO_OBJS \
    -L"$LIB" -lnvtt -lnvtt_b200 -Wl,-rpath,"\$ORIGIN/../../nvidia-texture-tools_b200/lib" -lpthread -ldl -o "$OUT/nvcompress_b200"
# the reference's
g++ -std=c++11 -O2 -w $REFINC "$REF/src/nvtt/tools/compress.cpp" $ALL_OBJS -lpthread -ldl -o "$OUT/nvcompress_ref"
echo "built $OUT/nvcompress_b200 $OUT/nvcompress_ref"
