"""INTEGRATION.md section B, compiled and run: integration/B200Compressor.h (the nv::CompressorInterface a maintainer adds to the
reference tree) is built against the reference's internal headers (tests/build_integration.sh) and driven with the reference's own
CompressionOptions / OutputOptions objects; its output must equal what the reference's CPU compressors write for the level."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_b200_compressor_binding(nvtt, ref):
    so = os.path.join(HERE, "_build", "libintegration_b200.so")
    if not os.path.exists(so):
        pytest.skip("tests/_build/libintegration_b200.so not built (needs /root/reference at build time)")
    L = C.CDLL(so)
    L.integ_compress_level.restype = C.c_long
    L.integ_compress_level.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_long]
    img = nvtt.synth.planar_from_bgra8(nvtt.synth.photo_bgra8(100, 60, seed=17, alpha=True))
    for fmt, q in ((ref.Format_BC1, 1), (ref.Format_BC1, 2), (ref.Format_BC3, 1), (ref.Format_BC5, 1), (ref.Format_BC4, 2), (ref.Format_BC7, 1), (ref.Format_BC6, 1)):
        cap = ref.level_size(fmt, 100, 60)
        out = np.empty(cap, np.uint8)
        n = L.integ_compress_level(fmt, q, 0, 100, 60, img.ctypes.data, 0, out.ctypes.data, cap)
        assert n == cap, (fmt, q, n)
        want = ref.compress_level(fmt, q, img)  # pixel type left at its default on both sides (BC6: unsigned)
        assert np.array_equal(out, want), (fmt, q)
    # Format_RGBA through the same binding (PixelFormatConverter): BGRA8 and R5G6B5
    for pf, masks, sizes, bpp in ((1, (32, 0xFF0000, 0xFF00, 0xFF, 0xFF000000), None, 4), (2, None, (5, 6, 5, 0), 2)):
        cap = 100 * 60 * bpp
        out = np.empty(cap, np.uint8)
        n = L.integ_compress_level(ref.Format_RGBA, 1, 0, 100, 60, img.ctypes.data, pf, out.ctypes.data, cap)
        assert n == cap
        bgra = nvtt.synth.photo_bgra8(100, 60, seed=17, alpha=True)
        want = ref.process([bgra], 0, 100, 60, ref.Format_RGBA, 1, mipmaps=False, gamma=(1.0, 1.0), pixel_masks=masks, pixel_sizes=sizes)
        assert np.array_equal(out, want), pf
