"""GPU parity against the committed golden fixtures (reference outputs) and against the plain-C oracle restatement."""
import hashlib
import os

import numpy as np
import pytest

import golden_cases as G

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def test_gpu_levels_match_golden(nvtt, ctx, golden):
    for key, (kind, w, h, fmt, q, am, cw, pt) in G.level_cases().items():
        img = G.make_input(kind, w, h, planar=True)
        got = ctx.encode_level(fmt, q, img, alpha_mode=am, color_weights=cw, pixel_type=pt)
        assert np.array_equal(got, golden[key]), key


def test_gpu_pipeline_matches_golden(nvtt, ctx, golden):
    for key, (kind, w, h, fmt, q, kw) in G.pipeline_cases().items():
        img = G.make_input(kind, w, h, planar=False)
        d = nvtt.make_process_desc(0, w, h, fmt, q, **kw)
        assert np.array_equal(ctx.process_bytes([img], d), golden[key]), key


def test_gpu_image_ops_match_golden(nvtt, ctx, golden):
    for key, (kind, w, h, wrap, filt, params) in G.imageop_cases().items():
        img = G.make_input(kind, w, h, planar=False)
        s = nvtt.Surface(ctx, wrap=wrap)
        s.set_image(0, w, h, img)
        s.to_linear(2.2)
        hashes = [_sha(s.get())]
        while s.build_next_mipmap(filt, params):
            hashes.append(_sha(s.get()))
        s.to_gamma(2.2)
        hashes.append(_sha(s.get()))
        assert np.array_equal(np.stack(hashes), golden[key]), key


@pytest.fixture(scope="module")
def golden2():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v2.npz"))


def test_gpu_v2_levels_decoders_quantize_metrics_match_golden(nvtt, ctx, golden, golden2):
    """golden_v2.npz (reference outputs, tests/golden/make_golden.py): the later format / quality pairs, every decoder
    flavour, quantize / binarize with and without Floyd-Steinberg, rmsError / rmsAlphaError."""
    lc1, lc2 = G.level_cases(), G.level_cases_v2()
    for key, (kind, w, h, fmt, q, am, cw, pt) in lc2.items():
        img = G.make_input(kind, w, h, planar=True)
        got = ctx.encode_level(fmt, q, img, alpha_mode=am, color_weights=cw, pixel_type=pt)
        assert np.array_equal(got, golden2[key]), key
    for key, (src, in_v2, dec) in G.decode_cases().items():
        kind, w, h, fmt = (lc2 if in_v2 else lc1)[src][:4]
        blocks = (golden2 if in_v2 else golden)[src]
        s = nvtt.Surface(ctx)
        s.set_image_2d(fmt, w, h, blocks, decoder=dec)
        assert np.array_equal(_sha(s.get()), golden2[key]), key
    for key, (w, h, dither) in G.quantize_cases().items():
        s = nvtt.Surface(ctx)
        s.set_image(0, w, h, nvtt.synth.photo_bgra8(w, h, seed=w + h, alpha=True))
        for ch, bits, exact in ((0, 5, 1), (1, 6, 1), (2, 3, 0)):
            ctx._ck(ctx.L.nvttb_surface_quantize(s.h, ch, bits, exact, dither))
        ctx._ck(ctx.L.nvttb_surface_binarize(s.h, 3, 0.4, dither))
        assert np.array_equal(_sha(s.get()), golden2[key]), key
    for name, src in (("bc1", "level_bc1_photo_48x40_q1"), ("bc3", "level_bc3_photo_48x40_q2")):
        kind, w, h, fmt = lc1[src][:4]
        for am in (0, 1):
            ref_s = nvtt.Surface(ctx, alpha_mode=am)
            ref_s.set_image(nvtt.InputFormat_RGBA_32F, w, h, np.ascontiguousarray(np.moveaxis(G.make_input(kind, w, h, planar=True), 0, 2)))
            dec = nvtt.Surface(ctx)
            dec.set_image_2d(fmt, w, h, golden[src])
            want = golden2["metric_%s_am%d" % (name, am)]
            got = (ref_s.rms_error(dec), ref_s.rms_alpha_error(dec))
            for g, wv in zip(got, want):
                assert abs(g - wv) <= 1e-6 * max(abs(float(wv)), 1e-30), (name, am, got, want)  # fp64 sums in a different order


def test_gpu_v2_containers_and_quantisation_pipeline_match_golden(nvtt, golden2):
    """KTX / DDS10 streams, colour dithering and binary alpha through the C++ mirror of nvtt::Compressor::process."""
    import ctypes as C
    import refapi
    so = os.path.join(ROOT, "tests", "_build", "libnvtt_b200_harness.so")
    if not os.path.exists(so):
        pytest.fail("tests/_build/libnvtt_b200_harness.so missing: run __graft_entry__.build()")
    L = C.CDLL(so)
    L.ref_process.restype = C.c_long
    L.ref_process.argtypes = [C.POINTER(refapi.RefProcessDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_long]
    for key, (kind, w, h, fmt, q, kw) in G.pipeline_cases_v2().items():
        img = np.ascontiguousarray(G.make_input(kind, w, h, planar=False))
        d = refapi.RefProcessDesc()
        d.inputFormat, d.textureType, d.width, d.height, d.faces = 0, 0, w, h, 1
        d.wrapMode, d.mipmapFilter, d.generateMipmaps, d.maxLevel = 2, kw.get("mip_filter", 0), 1, -1
        d.kaiserWidth, d.kaiserAlpha, d.kaiserStretch = 3.0, 4.0, 1.0
        d.inputGamma, d.outputGamma = 2.2, 2.2
        d.isNormalMap, d.convertToNormalMap, d.normalizeMipmaps = int(kw.get("normal_map", False)), 0, 1
        d.alphaMode, d.format, d.quality, d.pixelType = 0, fmt, q, 0
        d.colorWeights = (C.c_float * 4)(1, 1, 1, 1)
        d.outputHeader, d.container, d.threads = int(kw.get("header", False)), kw.get("container", 0), 0
        d.quantization, d.alphaThreshold = kw.get("quantization", 0), kw.get("alpha_threshold", 127)
        ptrs = (C.c_void_p * 1)(img.ctypes.data)
        n = L.ref_process(C.byref(d), ptrs, None, 0)
        assert n == golden2[key].size, (key, n, golden2[key].size)
        out = np.empty(n, np.uint8)
        assert L.ref_process(C.byref(d), ptrs, out.ctypes.data, n) == n
        assert np.array_equal(out, golden2[key]), key


def test_gpu_matches_oracle_restatement(nvtt, ctx):
    import oracleapi
    if not oracleapi.available():
        pytest.fail("oracle/_build/liboracle.so missing: run __graft_entry__.build()")
    s = nvtt.synth
    for (w, h) in ((128, 128), (61, 35)):
        col = s.photo_bgra8(w, h, seed=21, alpha=True)
        nrm = s.normal_bgra8(w, h, seed=22)
        for fmt, q, img, kw in ((1, 1, col, dict(mip_filter=0)), (1, 2, col, dict(mip_filter=2, alpha_mode=1)),
                                (4, 1, col, dict(mip_filter=2)), (7, 1, nrm, dict(mip_filter=2, normal_map=True)),
                                (6, 1, col, dict(mip_filter=1, wrap=0))):
            d = nvtt.make_process_desc(0, w, h, fmt, q, **kw)
            got = ctx.process_bytes([img], d)
            want = oracleapi.process([img], 0, w, h, fmt, q, **kw)
            assert np.array_equal(got, want), (w, h, fmt, q, kw)


def test_gpu_pixel_formats_match_golden_and_oracle(nvtt, ctx):
    """Format_RGBA writer (nvttb_convert_level) against golden_v3.npz (reference output) for 20 layouts x 2 sizes - one of
    which takes the four-pixels-per-thread kernel - and against the oracle on HDR / negative / Inf / NaN / denormal values
    and on a 1024 x 512 image (x4 kernel at full width, odd-width variant through the per-pixel kernel)."""
    import oracleapi
    if not oracleapi.available():
        pytest.fail("oracle/_build/liboracle.so missing: run __graft_entry__.build()")
    g3 = np.load(os.path.join(ROOT, "tests", "golden", "golden_v3.npz"))
    s = nvtt.synth
    for key, (w, h, kw) in G.pixel_format_cases().items():
        img = s.planar_from_bgra8(s.photo_bgra8(w, h, seed=w * 3 + h, alpha=True))
        got = ctx.convert_level(img, **kw)
        assert got.size == g3[key].size and np.array_equal(got, g3[key]), key
    rng = np.random.default_rng(11)
    for (w, h) in ((29, 7), (64, 9), (1024, 512), (1023, 3)):
        vals = np.exp(rng.normal(0, 4, (4, h, w))).astype(np.float32) * rng.choice([1.0, 1.0, 1.0, -1.0], (4, h, w)).astype(np.float32)
        vals[:, 0, :8] = np.array([0.0, -0.0, 1.0, 65504.0, 65520.0, 1e-8, 6.1e-5, 70000.0], np.float32)
        vals[:, h - 1, :4] = np.array([np.inf, -np.inf, np.nan, 5.96e-8], np.float32)
        unit = rng.random((4, h, w), dtype=np.float32) * 1.2 - 0.1
        for kw in (dict(sizes=(16, 16, 16, 16), pixel_type=4), dict(sizes=(32, 32, 32, 32), pixel_type=4), dict(sizes=(16, 0, 0, 0), pixel_type=4),
                   dict(sizes=(11, 11, 10, 0), pixel_type=4), dict(sizes=(8, 8, 8, 8), pixel_type=2), dict(sizes=(16, 16, 0, 0), pixel_type=2),
                   dict(), dict(masks=(16, 0xF800, 0x7E0, 0x1F, 0)), dict(masks=(8, 0xFF, 0, 0, 0)), dict(sizes=(10, 10, 10, 2)),
                   dict(masks=(24, 0xFF0000, 0xFF00, 0xFF, 0), pitch_alignment=4), dict(sizes=(3, 2, 2, 0))):
            for img in (vals, unit):
                got = ctx.convert_level(img, **kw)
                want = oracleapi.convert_level(img, **kw)
                assert np.array_equal(got, want), (w, h, kw)


@pytest.mark.gpu
def test_gpu_rgb9e5_matches_reference(nvtt, ctx, ref):
    """PixelType_SharedExp 9/9/9/5 (toFloat3SE, CompressorRGB.cpp:231-269) on the GPU against the reference: magnitudes across
    the range, specials, and the floats either side of every power of two (floor(log2f) rounding); by sizes and by masks."""
    from test_oracle import rgb9e5_probe_values
    vals, w, h = rgb9e5_probe_values()
    planar = np.ascontiguousarray(np.moveaxis(vals, 2, 0))
    for kw in (dict(sizes=(9, 9, 9, 5)), dict(masks=(32, 0x1FF, 0x3FE00, 0x7FC0000, 0xF8000000))):
        got = ctx.convert_level(planar, pixel_type=6, **kw)
        want = ref.process([vals], 2, w, h, 0, 1, mipmaps=False, gamma=(1.0, 1.0), pixel_masks=kw.get("masks"), pixel_sizes=kw.get("sizes"),
                           pixel_type=6)
        assert got.size == want.size and np.array_equal(got, want), (kw, int(np.flatnonzero(got != want)[0]) // 4)
