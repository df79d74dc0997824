"""CPU tests of the oracle (oracle/oracle.c, plain-C restatement): pinned against the golden vectors generated from the
unmodified reference (tests/golden/golden_v1.npz) and, where the reference build is present, against it directly."""
import hashlib
import os

import numpy as np
import pytest

import golden_cases as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def orc():
    import oracleapi
    if not oracleapi.available():
        pytest.skip("oracle/_build/liboracle.so not built (run __graft_entry__.build())")
    return oracleapi


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def test_oracle_levels_match_golden(orc, golden):
    for key, (kind, w, h, fmt, q, am, cw, pt) in G.level_cases().items():
        if fmt in (2, 3, 5, 10, 11) or (fmt == 4 and q == 0):
            continue  # BC2 / BC3n / BC6H / BC7 are checked against the compiled reference itself (oracle/_ref), not restated in oracle.c
        img = G.make_input(kind, w, h, planar=True)
        got = orc.compress_level(fmt, q, img, am, cw)
        assert np.array_equal(got, golden[key]), key


def test_oracle_pipeline_matches_golden(orc, golden):
    for key, (kind, w, h, fmt, q, kw) in G.pipeline_cases().items():
        img = G.make_input(kind, w, h, planar=False)
        got = orc.process([img], 0, w, h, fmt, q, **kw)
        assert np.array_equal(got, golden[key]), key


def test_oracle_image_ops_match_golden(orc, golden):
    for key, (kind, w, h, wrap, filt, params) in G.imageop_cases().items():
        img = G.make_input(kind, w, h, planar=False)
        cur = orc.to_linear(orc.set_image(0, w, h, img), 2.2)
        hashes = [_sha(cur)]
        fw, p0, p1 = (0.5, 0, 0) if filt == 0 else (1.0, 0, 0) if filt == 1 else params
        while cur.shape[1] > 1 or cur.shape[2] > 1:
            cur = orc.next_mipmap(cur, filt, fw, p0, p1, wrap)
            hashes.append(_sha(cur))
        hashes.append(_sha(orc.to_gamma(cur, 2.2)))
        assert np.array_equal(np.stack(hashes), golden[key]), key


def test_oracle_pixel_formats_match_golden(orc, nvtt):
    """Format_RGBA layouts: orc_convert_level against golden_v3.npz (the reference's Compressor::process output)."""
    g3 = np.load(os.path.join(ROOT, "tests", "golden", "golden_v3.npz"))
    s = nvtt.synth
    for key, (w, h, kw) in G.pixel_format_cases().items():
        img = s.planar_from_bgra8(s.photo_bgra8(w, h, seed=w * 3 + h, alpha=True))
        got = orc.convert_level(img, **kw)
        assert got.size == g3[key].size and np.array_equal(got, g3[key]), key


def test_oracle_pixel_formats_vs_reference_direct(orc, ref, nvtt):
    """HDR / out-of-range / special values through the float, uint and fixed layouts against the reference itself."""
    rng = np.random.default_rng(11)
    w, h = 29, 7
    vals = np.exp(rng.normal(0, 4, (h, w, 4))).astype(np.float32) * rng.choice([1.0, 1.0, 1.0, -1.0], (h, w, 4)).astype(np.float32)
    vals[0, :8, :] = np.array([0.0, -0.0, 1.0, 65504.0, 65520.0, 1e-8, 6.1e-5, 70000.0], np.float32)[:, None]
    vals[1, :4, :] = np.array([np.inf, -np.inf, np.nan, 5.96e-8], np.float32)[:, None]
    planar = np.ascontiguousarray(np.moveaxis(vals, 2, 0))
    for kw in (dict(sizes=(16, 16, 16, 16), pixel_type=4), dict(sizes=(32, 32, 32, 32), pixel_type=4), dict(sizes=(16, 0, 0, 0), pixel_type=4),
               dict(sizes=(11, 11, 10, 0), pixel_type=4), dict(sizes=(8, 8, 8, 8), pixel_type=2), dict(sizes=(16, 16, 0, 0), pixel_type=2),
               dict(), dict(masks=(16, 0xF800, 0x7E0, 0x1F, 0)), dict(sizes=(10, 10, 10, 2))):
        a = orc.convert_level(planar, **kw)
        b = ref.process([vals], 2, w, h, 0, 1, mipmaps=False, gamma=(1.0, 1.0), pixel_masks=kw.get("masks"), pixel_sizes=kw.get("sizes"),
                        pixel_type=kw.get("pixel_type", 0), pitch_alignment=kw.get("pitch_alignment", 0))
        assert np.array_equal(a, b), (kw, int(np.flatnonzero(a != b)[0]) if a.size == b.size else (a.size, b.size))


def rgb9e5_probe_values():
    """Inputs for the R9G9B9E5 writer: magnitudes across its range, specials, and the 48 floats either side of every power of
    two from 2^-20 to 2^17 (where floor(log2f(x)) depends on how log2f rounds)."""
    rng = np.random.default_rng(5)
    v = [np.exp(rng.normal(0, 5, 4000)).astype(np.float32), rng.random(1000, dtype=np.float32) * 70000.0,
         np.array([0.0, -0.0, 1e-30, 1e-8, 255.5, 256.0, 65408.0, 65408.01, 70000.0, np.inf, -np.inf, np.nan, -1.0, 5.9e-39], np.float32)]
    for k in range(-20, 18):
        base = np.float32(2.0 ** k).view(np.uint32)
        v.append((np.arange(-48, 48, dtype=np.int64) + int(base)).astype(np.uint32).view(np.float32))
    v = np.concatenate(v)
    n = (v.size + 2) // 3 * 3
    v = np.concatenate([v, np.zeros(n - v.size, np.float32)])
    rgb = np.stack([v, np.roll(v, 1) * np.float32(0.37), np.roll(v, 2) * np.float32(0.051)], 1)  # every value is the maximum somewhere
    w = 64
    h = (rgb.shape[0] + w - 1) // w
    img = np.zeros((h * w, 4), np.float32)
    img[:rgb.shape[0], :3] = rgb
    img[:, 3] = 1.0
    return img.reshape(h, w, 4), w, h


def test_oracle_rgb9e5_vs_reference(orc, ref):
    """PixelType_SharedExp 9/9/9/5 (toFloat3SE, CompressorRGB.cpp:231-269), by sizes and by masks, against the reference itself -
    including its x86 shift-count behaviour (zero mantissas below 256) and the log2f rounding at powers of two."""
    vals, w, h = rgb9e5_probe_values()
    planar = np.ascontiguousarray(np.moveaxis(vals, 2, 0))
    for kw in (dict(sizes=(9, 9, 9, 5)), dict(masks=(32, 0x1FF, 0x3FE00, 0x7FC0000, 0xF8000000))):
        a = orc.convert_level(planar, pixel_type=6, **kw)
        b = ref.process([vals], 2, w, h, 0, 1, mipmaps=False, gamma=(1.0, 1.0), pixel_masks=kw.get("masks"), pixel_sizes=kw.get("sizes"),
                        pixel_type=6)
        assert a.size == b.size and np.array_equal(a, b), (kw, int(np.flatnonzero(a != b)[0]) // 4)
    assert np.count_nonzero(b.view(np.uint32) & 0x7FFFFFF) > 1000  # the mantissas are not all zero: values >= 256 are encoded


def test_oracle_vs_reference_direct(orc, ref, nvtt):
    """Wider sweep against the reference itself (only where oracle/_ref exists)."""
    s = nvtt.synth
    rng = np.random.default_rng(3)
    for (w, h) in ((32, 32), (21, 9), (4, 4), (2, 3), (1, 1)):
        imgs = [s.planar_from_bgra8(s.photo_bgra8(w, h, seed=8, alpha=True)), s.planar_from_bgra8(s.adversarial_bgra8(w, h, seed=2)),
                rng.random((4, h, w), dtype=np.float32) * 1.5 - 0.25]
        for img in imgs:
            for fmt, qs in ((1, (0, 1, 2, 3)), (4, (1, 2, 3)), (6, (0, 1, 2)), (7, (0, 1, 3))):
                for q in qs:
                    for am in (0, 1):
                        a = orc.compress_level(fmt, q, img, am, (0.8, 1.0, 0.6, 1.0))
                        b = ref.compress_level(fmt, q, img, alpha_mode=am, color_weights=(0.8, 1.0, 0.6, 1.0))
                        assert np.array_equal(a, b), (w, h, fmt, q, am)
    # every half bit pattern through setImage (half_to_float, Half.cpp)
    allh = np.arange(65536, dtype=np.uint16).reshape(256, 256, 1)
    sweep = np.concatenate([allh, allh[::-1], np.roll(allh, 7, 0), allh.transpose(1, 0, 2)], 2).copy()
    r = ref.Surface()
    r.set_image(1, 256, 256, sweep)
    assert np.array_equal(orc.set_image(1, 256, 256, sweep).view(np.uint32), r.get().view(np.uint32))
    # fp16 input + Mitchell resize + renormalise
    hh = rng.integers(0, 65536, (16, 24, 4)).astype(np.uint16)
    r = ref.Surface()
    r.set_image(1, 24, 16, hh)
    assert np.array_equal(orc.set_image(1, 24, 16, hh).view(np.uint32), r.get().view(np.uint32))
    im8 = s.normal_bgra8(40, 24)
    r = ref.Surface(wrap=1, normal_map=True)
    r.set_image(0, 40, 24, im8)
    r.resize(17, 11, 3)
    o = orc.resize(orc.set_image(0, 40, 24, im8), 17, 11, 3, 2.0, 1.0 / 3.0, 1.0 / 3.0, 1)
    assert np.array_equal(o.view(np.uint32), r.get().view(np.uint32))
    r.expand_normals(); r.normalize_normal_map(); r.pack_normals()
    assert np.array_equal(orc.renormalize(o).view(np.uint32), r.get().view(np.uint32))


def test_oracle_dense_sweeps_vs_reference(orc, ref):
    """The oracle's gamma approximations and half_from_float against the reference on dense sweeps of the fp32 bit patterns:
    every 256th pattern of [0, 1] through toLinear / toGamma (4 M values per channel, three residues) and every 1021st pattern of
    the whole 32-bit range through the RGBA16F writer (4 M values)."""
    n = 1024
    base = np.arange(n * n, dtype=np.uint64) * 1016  # 1024 * 1024 * 1016 ~ 0x3F800000
    img = np.empty((n, n, 4), np.uint32)
    for c, off in enumerate((0, 337, 674)):
        img[..., c] = np.minimum(base + off, np.uint64(0x3F800000)).astype(np.uint32).reshape(n, n)
    img[..., 3] = 0x3F800000
    special = np.array([1.0000001, 1.5, 255.0, 3.4e38, -0.0, -1e-30, -0.5, np.inf, -np.inf, np.nan, 1e-45], np.float32).view(np.uint32)
    img[0, :special.size, 0] = special
    data = img.view(np.float32)
    for op in ("to_linear", "to_gamma"):
        r = ref.Surface()
        r.set_image(2, n, n, data)
        getattr(r, op)(2.2)
        o = getattr(orc, op)(orc.set_image(2, n, n, data), 2.2)
        assert np.array_equal(o.view(np.uint32), r.get().view(np.uint32)), op
    w, h = 1024, 1024
    bits = (np.arange(w * h * 4, dtype=np.uint64) * 1021 % (1 << 32)).astype(np.uint32).reshape(h, w, 4)
    vals = bits.view(np.float32)
    a = orc.convert_level(np.ascontiguousarray(np.moveaxis(vals, 2, 0)), sizes=(16, 16, 16, 16), pixel_type=4)
    b = ref.process([vals], 2, w, h, 0, 1, mipmaps=False, gamma=(1.0, 1.0), pixel_sizes=(16, 16, 16, 16), pixel_type=4)
    assert np.array_equal(a, b)
