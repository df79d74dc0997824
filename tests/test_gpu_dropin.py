"""Drop-in test of the C++ boundary: oracle/ref_harness.cpp — a caller written against the REFERENCE's nvtt.h — is
compiled unchanged against nvidia-texture-tools_b200/host/nvtt/nvtt.h and linked with our libnvtt.so.  The identical
nvtt:: call sequences must give byte-identical output (DDS header included) to the reference library."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ours(ref):
    so = os.path.join(ROOT, "tests", "_build", "libnvtt_b200_harness.so")
    if not os.path.exists(so):
        pytest.fail("tests/_build/libnvtt_b200_harness.so missing: run __graft_entry__.build()")
    return ref._load(so) if os.path.isabs(so) and False else _load_harness(ref, so)


def _load_harness(ref, so):
    # same prototypes as the reference harness, different library
    import types
    L = C.CDLL(so)
    L.ref_compress_level.restype = C.c_long
    L.ref_compress_level.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_long]
    L.ref_process.restype = C.c_long
    L.ref_process.argtypes = [C.POINTER(ref.RefProcessDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_long]
    L.ref_surf_create.restype = C.c_void_p
    L.ref_surf_create.argtypes = [C.c_int, C.c_int, C.c_int]
    L.ref_surf_destroy.argtypes = [C.c_void_p]
    L.ref_surf_set_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.ref_surf_width.argtypes = [C.c_void_p]
    L.ref_surf_height.argtypes = [C.c_void_p]
    L.ref_surf_get.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_surf_to_linear.argtypes = [C.c_void_p, C.c_float]
    L.ref_surf_to_gamma.argtypes = [C.c_void_p, C.c_float]
    L.ref_surf_build_next_mipmap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    L.ref_surf_quantize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ref_surf_binarize.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
    return L


def _process(L, ref, images, fmt, quality, w, h, **kw):
    d = ref.RefProcessDesc()
    d.inputFormat, d.textureType, d.width, d.height, d.faces = 0, kw.get("texture_type", 0), w, h, len(images)
    d.wrapMode, d.mipmapFilter, d.generateMipmaps, d.maxLevel = kw.get("wrap", 2), kw.get("mip_filter", 0), 1, -1
    d.kaiserWidth, d.kaiserAlpha, d.kaiserStretch = 3.0, 4.0, 1.0
    d.inputGamma, d.outputGamma = 2.2, 2.2
    d.isNormalMap, d.convertToNormalMap, d.normalizeMipmaps = int(kw.get("normal_map", False)), 0, 1
    d.alphaMode, d.format, d.quality, d.pixelType = kw.get("alpha_mode", 0), fmt, quality, kw.get("pixel_type", 0)
    d.inputFormat = kw.get("input_format", 0)
    ref.set_pixel_format(d, kw.get("pixel_masks"), kw.get("pixel_sizes"), kw.get("pitch_alignment", 0))
    d.colorWeights = (C.c_float * 4)(1, 1, 1, 1)
    d.outputHeader, d.container, d.threads = int(kw.get("header", True)), kw.get("container", 0), 0
    d.quantization, d.alphaThreshold = kw.get("quantization", 0), kw.get("alpha_threshold", 127)
    imgs = [np.ascontiguousarray(i) for i in images]
    ptrs = (C.c_void_p * len(imgs))(*[i.ctypes.data for i in imgs])
    n = L.ref_process(C.byref(d), ptrs, None, 0)
    assert n > 0, "process failed"
    out = np.empty(n, np.uint8)
    assert L.ref_process(C.byref(d), ptrs, out.ctypes.data, n) == n
    return out


def test_process_with_header_identical(nvtt, ref, ours):
    s = nvtt.synth
    cases = [
        (ref.Format_BC1, 1, s.photo_bgra8(128, 64, seed=1), dict(mip_filter=0)),
        (ref.Format_BC3, 1, s.photo_bgra8(128, 128, seed=2, alpha=True), dict(mip_filter=2)),
        (ref.Format_BC5, 1, s.normal_bgra8(64, 64), dict(mip_filter=2, normal_map=True)),
        (ref.Format_BC4, 1, s.photo_bgra8(37, 22, seed=3), dict(mip_filter=1, container=1)),
        (ref.Format_BC7, 1, s.photo_bgra8(16, 12, seed=4, alpha=True), dict(mip_filter=0, container=1)),
        # KTX: mip-major order with size words; Kaiser parameters swapped and pack-less renormalisation (reference quirks)
        (ref.Format_BC1, 1, s.photo_bgra8(64, 32, seed=7), dict(mip_filter=0, container=2)),
        (ref.Format_BC3, 1, s.photo_bgra8(40, 40, seed=8, alpha=True), dict(mip_filter=2, container=2)),
        (ref.Format_BC5, 1, s.normal_bgra8(32, 32), dict(mip_filter=1, normal_map=True, container=2)),
    ]
    for fmt, q, img, kw in cases:
        h, w = img.shape[:2]
        got = _process(ours, ref, [img], fmt, q, w, h, **kw)
        want = _process(ref.lib(), ref, [img], fmt, q, w, h, **kw)
        assert got.size == want.size
        assert np.array_equal(got, want), (fmt, kw)


def test_quantization_settings_identical(nvtt, ref, ours):
    """CompressionOptions::setQuantization as nvcompress uses it (-bc1a: alpha dithering + binary alpha, -bc2: alpha
    dithering: both no-ops for BCn, Context.cpp:519-541) and binary alpha alone (non-dithered binarize of the alpha plane)."""
    img = nvtt.synth.photo_bgra8(64, 48, seed=31, alpha=True)
    for fmt, q in ((ref.Format_DXT1a, 2 | 4), (ref.Format_BC2, 2), (ref.Format_DXT1a, 4), (ref.Format_BC3, 4)):
        for thr in (127, 40):
            kw = dict(mip_filter=0, quantization=q, alpha_threshold=thr)
            got = _process(ours, ref, [img], fmt, 1, 64, 48, **kw)
            want = _process(ref.lib(), ref, [img], fmt, 1, 64, 48, **kw)
            assert got.size == want.size and np.array_equal(got, want), (fmt, q, thr)


def test_surface_quantize_and_binarize_identical(nvtt, ref, ours):
    """Surface::quantize / binarize, plain and with the Floyd-Steinberg scan (more rows than one 1024-row band, odd sizes)."""
    ref.lib().ref_surf_quantize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    ref.lib().ref_surf_binarize.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
    for (w, h) in ((37, 22), (64, 1100), (1, 5), (300, 3)):
        im8 = nvtt.synth.photo_bgra8(w, h, seed=w + h, alpha=True)
        for dither in (0, 1):
            outs = []
            for L in (ours, ref.lib()):
                s = L.ref_surf_create(2, 0, 0)
                assert L.ref_surf_set_image(s, 0, w, h, im8.ctypes.data)
                L.ref_surf_quantize(s, 0, 5, 1, dither)
                L.ref_surf_quantize(s, 1, 6, 1, dither)
                L.ref_surf_quantize(s, 2, 3, 0, dither)
                L.ref_surf_binarize(s, 3, 0.4, dither)
                o = np.empty((4, h, w), np.float32)
                L.ref_surf_get(s, o.ctypes.data)
                L.ref_surf_destroy(s)
                outs.append(o)
            assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32)), (w, h, dither)


def test_color_dithering_pipeline_identical(nvtt, ref, ours):
    """setQuantization(colorDithering = true): 5/6/5-bit Floyd-Steinberg on every mip level before BC1 / BC3 (Context.cpp:519-531)."""
    img = nvtt.synth.photo_bgra8(96, 72, seed=41, alpha=True)
    for fmt in (ref.Format_BC1, ref.Format_BC3):
        kw = dict(mip_filter=0, quantization=1)
        got = _process(ours, ref, [img], fmt, 1, 96, 72, **kw)
        want = _process(ref.lib(), ref, [img], fmt, 1, 96, 72, **kw)
        assert got.size == want.size and np.array_equal(got, want), fmt


def test_process_general_path_extents_round_modes_user_mips(nvtt, ref, ours):
    """Compressor::process outside the fused pipeline (Context.cpp:233-245,291-323; Surface.cpp:220-327 getTargetExtent): input
    resize through setMaxExtents with every RoundMode, and caller-supplied mip levels (all of them, or only the first ones -
    the chain continues from the last supplied level with the filter)."""
    img = nvtt.synth.photo_bgra8(200, 120, seed=12, alpha=True)
    for fmt, q, kw in ((ref.Format_BC1, 1, dict(mip_filter=0)), (ref.Format_BC3, 1, dict(mip_filter=2)), (ref.Format_BC5, 1, dict(mip_filter=1, normal_map=True))):
        for round_mode in range(7):
            for max_extent in (0, 64, 100):
                if round_mode == 0 and max_extent == 0:
                    continue
                a = ref.process([img], 0, 200, 120, fmt, q, header=True, max_extent=max_extent, round_mode=round_mode, lib_override=ours, **kw)
                b = ref.process([img], 0, 200, 120, fmt, q, header=True, max_extent=max_extent, round_mode=round_mode, **kw)
                assert a.size == b.size and np.array_equal(a, b), (fmt, round_mode, max_extent)
    # user-supplied mip levels
    base = nvtt.synth.photo_bgra8(64, 32, seed=13, alpha=True)
    mips, w, h, m = {}, 64, 32, 0
    while w > 1 or h > 1:
        w, h, m = max(1, w // 2), max(1, h // 2), m + 1
        mips[(0, m)] = nvtt.synth.photo_bgra8(w, h, seed=100 + m, alpha=True)  # deliberately unrelated to level 0
    first_two = {k: v for k, v in mips.items() if k[1] <= 2}
    for fmt, kw in ((ref.Format_BC1, dict(mip_filter=0)), (ref.Format_BC3, dict(mip_filter=2)), (ref.Format_BC1, dict(mip_filter=0, container=2))):
        for user in (mips, first_two):
            a = ref.process([base], 0, 64, 32, fmt, 1, header=True, user_mips=user, lib_override=ours, **kw)
            b = ref.process([base], 0, 64, 32, fmt, 1, header=True, user_mips=user, **kw)
            assert a.size == b.size and np.array_equal(a, b), (fmt, kw, len(user))


def test_ktx_cube_identical(nvtt, ref, ours):
    """Six faces through the KTX container: the faces of a level are stored together (Context.cpp:347-472)."""
    faces = [nvtt.synth.photo_bgra8(32, 32, seed=20 + f) for f in range(6)]
    got = _process(ours, ref, faces, ref.Format_BC1, 1, 32, 32, mip_filter=0, container=2, texture_type=1)
    want = _process(ref.lib(), ref, faces, ref.Format_BC1, 1, 32, 32, mip_filter=0, container=2, texture_type=1)
    assert got.size == want.size and np.array_equal(got, want)


def test_raw_compress_and_surface_api_identical(nvtt, ref, ours):
    img = nvtt.synth.planar_from_bgra8(nvtt.synth.photo_bgra8(64, 64, seed=5, alpha=True))
    for fmt in (ref.Format_BC1, ref.Format_BC3, ref.Format_BC4, ref.Format_BC5):
        for am in (0, 1):
            n = ref.level_size(fmt, 64, 64)
            a = np.empty(n, np.uint8)
            b = np.empty(n, np.uint8)
            assert ours.ref_compress_level(fmt, 1, am, 64, 64, img.ctypes.data, None, 0, 0, a.ctypes.data, n) == n
            assert ref.lib().ref_compress_level(fmt, 1, am, 64, 64, img.ctypes.data, None, 0, 0, b.ctypes.data, n) == n
            assert np.array_equal(a, b), (fmt, am)
    # imperative Surface walk-through (tests/imperativeapi.cpp of the reference): toLinear, Kaiser mips, toGamma
    im8 = nvtt.synth.photo_bgra8(96, 40, seed=6)
    outs = []
    for L in (ours, ref.lib()):
        h = L.ref_surf_create(2, 0, 0)
        assert L.ref_surf_set_image(h, 0, 96, 40, im8.ctypes.data)
        L.ref_surf_to_linear(h, 2.2)
        levels = []
        while L.ref_surf_build_next_mipmap(h, 2, 1, 3.0, 4.0, 1.0):
            o = np.empty((4, L.ref_surf_height(h), L.ref_surf_width(h)), np.float32)
            L.ref_surf_get(h, o.ctypes.data)
            levels.append(o)
        L.ref_surf_destroy(h)
        outs.append(levels)
    assert len(outs[0]) == len(outs[1]) == 6
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_rgba_pixel_formats_identical(nvtt, ref, ours):
    """Format_RGBA (PixelFormatConverter, CompressorRGB.cpp:410-575) through Compressor::process: default BGRA8, mask and size
    layouts, luminance, 16-bit, 24-bit with 4-byte pitch alignment, half / float / 11-11-10 float channels, UnsignedInt,
    the zero-filled signed types, odd widths; DDS header included."""
    s = nvtt.synth
    w, h = 37, 22
    img = s.photo_bgra8(w, h, seed=21, alpha=True)
    cases = [
        dict(),
        dict(pixel_masks=(32, 0xFF, 0xFF00, 0xFF0000, 0xFF000000)),
        dict(pixel_masks=(24, 0xFF0000, 0xFF00, 0xFF, 0), pitch_alignment=4),
        dict(pixel_masks=(16, 0xF800, 0x7E0, 0x1F, 0)),
        dict(pixel_masks=(16, 0x7C00, 0x3E0, 0x1F, 0x8000), pitch_alignment=8),
        dict(pixel_masks=(8, 0xFF, 0, 0, 0)),
        dict(pixel_masks=(8, 0, 0, 0, 0xFF)),
        dict(pixel_masks=(32, 0x3FF00000, 0xFFC00, 0x3FF, 0xC0000000)),
        dict(pixel_masks=(12, 0xF00, 0xF0, 0xF, 0)),
        dict(pixel_sizes=(5, 6, 5, 0)),
        dict(pixel_sizes=(8, 8, 8, 8)),
        dict(pixel_sizes=(10, 10, 10, 2)),
        dict(pixel_sizes=(16, 0, 0, 0)),
        dict(pixel_sizes=(3, 3, 2, 0), quantization=1),
        dict(pixel_sizes=(4, 4, 4, 4), quantization=3),
        dict(pixel_sizes=(8, 8, 8, 8), pixel_type=2),
        dict(pixel_sizes=(8, 8, 8, 8), pixel_type=1),
        dict(pixel_sizes=(16, 16, 16, 16), pixel_type=4),
        dict(pixel_sizes=(32, 32, 32, 32), pixel_type=4),
        dict(pixel_sizes=(16, 16, 0, 0), pixel_type=4),
        dict(pixel_sizes=(32, 0, 0, 0), pixel_type=4, pitch_alignment=4),
        dict(pixel_sizes=(11, 11, 10, 0), pixel_type=4, header=False),
        dict(pixel_sizes=(9, 9, 9, 5), pixel_type=6),   # R9G9B9E5 (toFloat3SE)
        dict(pixel_masks=(32, 0x1FF, 0x3FE00, 0x7FC0000, 0xF8000000), pixel_type=6),
        dict(pixel_sizes=(10, 10, 10, 2), pixel_type=6),  # other shared-exponent layouts: zeros (8/8/8/8 writes nothing at all in the
                                                          # reference - an empty "@@" branch - so its scanlines are uninitialised memory)
    ]
    for kw in cases:
        kw = dict(kw)
        kw.setdefault("header", True)
        got = _process(ours, ref, [img], ref.Format_RGBA, 1, w, h, mip_filter=0, **kw)
        want = _process(ref.lib(), ref, [img], ref.Format_RGBA, 1, w, h, mip_filter=0, **kw)
        assert got.size == want.size, kw
        assert np.array_equal(got, want), (kw, int(np.flatnonzero(got != want)[0]))
    # widths that are a multiple of four take the four-pixels-per-thread kernel
    img4 = s.photo_bgra8(64, 24, seed=22, alpha=True)
    for kw in (dict(), dict(pixel_masks=(16, 0xF800, 0x7E0, 0x1F, 0)), dict(pixel_masks=(8, 0xFF, 0, 0, 0)), dict(pixel_sizes=(10, 10, 10, 2)),
               dict(pixel_sizes=(16, 16, 16, 16), pixel_type=4), dict(pixel_sizes=(32, 32, 32, 32), pixel_type=4),
               dict(pixel_masks=(24, 0xFF0000, 0xFF00, 0xFF, 0)), dict(pixel_sizes=(8, 8, 8, 8), pitch_alignment=8)):
        got = _process(ours, ref, [img4], ref.Format_RGBA, 1, 64, 24, mip_filter=0, **kw)
        want = _process(ref.lib(), ref, [img4], ref.Format_RGBA, 1, 64, 24, mip_filter=0, **kw)
        assert np.array_equal(got, want), kw
    # KTX and DDS10 containers (glType / glFormat lookup, DXGI lookup)
    for kw in (dict(container=2), dict(container=2, pixel_masks=(32, 0xFF, 0xFF00, 0xFF0000, 0xFF000000)),
               dict(container=2, pixel_masks=(24, 0xFF0000, 0xFF00, 0xFF, 0)), dict(container=2, pixel_sizes=(16, 0, 0, 0)),
               dict(container=2, pixel_sizes=(16, 16, 16, 16), pixel_type=4), dict(container=1), dict(container=1, pixel_masks=(16, 0xF800, 0x7E0, 0x1F, 0)),
               dict(container=1, pixel_sizes=(16, 0, 0, 0)), dict(container=1, pixel_masks=(8, 0xFF, 0, 0, 0))):
        got = _process(ours, ref, [img], ref.Format_RGBA, 1, w, h, mip_filter=0, **kw)
        want = _process(ref.lib(), ref, [img], ref.Format_RGBA, 1, w, h, mip_filter=0, **kw)
        assert np.array_equal(got, want), kw
    # HDR input through the float layouts, DX10 container
    hdr = s.hdr_rgba16f(24, 16, seed=5).view("uint16")
    for kw in (dict(pixel_sizes=(16, 16, 16, 16), pixel_type=4, container=1), dict(pixel_sizes=(11, 11, 10, 0), pixel_type=4, container=1),
               dict(pixel_sizes=(32, 32, 32, 32), pixel_type=4)):
        got = _process(ours, ref, [hdr], ref.Format_RGBA, 1, 24, 16, mip_filter=0, input_format=1, **kw)
        want = _process(ref.lib(), ref, [hdr], ref.Format_RGBA, 1, 24, 16, mip_filter=0, input_format=1, **kw)
        assert np.array_equal(got, want), kw
