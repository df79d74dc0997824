#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  The reference ships real test images (data/testsuite/kodak, id_tnmap, data/witness - SURVEY.md 8c);
# /root/reference does not exist on the GPU box, so the ones the GPU parity tests use are copied here (images, not sources)
# into tests/data/_ref_images/, which is git-ignored like oracle/_ref/ and travels to the box with the working tree.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${NVTT_REFERENCE:-/root/reference}"
if [ ! -d "$REF/data/testsuite/kodak" ]; then echo "reference data not found (existing tests/data/_ref_images is used as is)"; exit 0; fi
mkdir -p "$HERE/_ref_images/kodak" "$HERE/_ref_images/id_tnmap" "$HERE/_ref_images/witness"
cp -u "$REF"/data/testsuite/kodak/*.png "$HERE/_ref_images/kodak/"
cp -u "$REF"/data/testsuite/id_tnmap/*.png "$HERE/_ref_images/id_tnmap/" 2>/dev/null || true
cp -u "$REF"/data/testsuite/id_tnmap/*.tga "$HERE/_ref_images/id_tnmap/" 2>/dev/null || true
cp -u "$REF"/data/witness/*.dds "$HERE/_ref_images/witness/"
echo "copied test images into $HERE/_ref_images"
