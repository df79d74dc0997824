"""GPU parity tests: the CUDA path (through the C ABI) against the unmodified reference built with the pinned
parity flags (oracle/_ref) on the same seeded inputs.  Bar: bit-exact for every integer / block output and for the
fp32 image ops (they are implemented operation-for-operation without FMA contraction)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [(64, 64), (128, 32), (37, 22), (4, 4), (3, 2), (1, 1), (5, 9), (260, 4)]


def _images(nvtt, w, h):
    s = nvtt.synth
    yield "photo", s.planar_from_bgra8(s.photo_bgra8(w, h, seed=1234, alpha=True))
    yield "normal", s.planar_from_bgra8(s.normal_bgra8(w, h, seed=7))
    yield "adversarial", s.planar_from_bgra8(s.adversarial_bgra8(w, h, seed=5))
    rng = np.random.default_rng(w * 1000 + h)
    yield "float_oob", (rng.random((4, h, w), dtype=np.float32) * 1.5 - 0.25)  # exercises the clamp in the quantiser


def _assert_blocks_equal(got, want, bs, what):
    assert got.shape == want.shape, what
    bad = (got.reshape(-1, bs) != want.reshape(-1, bs)).any(1)
    if bad.any():
        i = int(np.nonzero(bad)[0][0])
        raise AssertionError("%s: %d/%d blocks differ; first block %d got %s want %s"
                             % (what, int(bad.sum()), bad.size, i, got.reshape(-1, bs)[i], want.reshape(-1, bs)[i]))


@pytest.mark.parametrize("fmt_name,quality", [("BC1", 0), ("BC1", 1), ("BC1", 2), ("BC1", 3), ("BC4", 0), ("BC4", 1), ("BC5", 0),
                                              ("BC5", 1), ("BC3", 1), ("BC3", 2), ("BC4", 2), ("BC5", 2), ("BC5", 3), ("BC3", 3), ("BC2", 1), ("BC2", 2), ("BC3n", 1), ("BC3n", 2), ("BC3n", 3), ("BC2", 0), ("BC3", 0), ("BC3n", 0)])
def test_level_encode_bit_exact(nvtt, ref, ctx, fmt_name, quality):
    fmt = getattr(nvtt, "Format_" + fmt_name)
    bs = 8 if fmt_name in ("BC4", "BC1") else 16
    for (w, h) in SIZES:
        for name, img in _images(nvtt, w, h):
            got = ctx.encode_level(fmt, quality, img)
            want = ref.compress_level(fmt, quality, img)
            _assert_blocks_equal(got, want, bs, "%s q%d %s %dx%d" % (fmt_name, quality, name, w, h))


def _hdr_planar(nvtt, w, h, seed):
    f = nvtt.synth.hdr_rgba16f(w, h, seed=seed).astype(np.float32)
    return np.ascontiguousarray(f.transpose(2, 0, 1))


def test_bc6h_level_bit_exact(nvtt, ref, ctx):
    """BC6H (ZOH): unsigned and signed half, HDR / LDR / out-of-range / flat inputs, ragged sizes, alpha-weighted texels."""
    rng = np.random.default_rng(17)
    for (w, h) in SIZES:
        wild = (rng.standard_normal((4, h, w)) * np.exp(rng.normal(0, 4, (4, h, w)))).astype(np.float32)
        imgs = [("hdr", _hdr_planar(nvtt, w, h, 11)), ("ldr", nvtt.synth.planar_from_bgra8(nvtt.synth.photo_bgra8(w, h, seed=2, alpha=True))),
                ("wild", wild), ("flat", np.full((4, h, w), 0.5, np.float32))]
        for name, img in imgs:
            for pt in (nvtt.PixelType_UnsignedFloat, nvtt.PixelType_Float):
                for am in (0, 1):
                    got = ctx.encode_level(nvtt.Format_BC6, 1, img, pixel_type=pt, alpha_mode=am)
                    want = ref.compress_level(ref.Format_BC6, 1, img, pixel_type=pt, alpha_mode=am)
                    _assert_blocks_equal(got, want, 16, "BC6H %s %dx%d pixelType %d alphaMode %d" % (name, w, h, pt, am))


def test_bc6h_cubemap_pipeline_bit_exact(nvtt, ref, ctx):
    """BASELINE configs[4] in small: 6 fp16 faces, gamma-correct (2.2 in / 2.2 out) and linear (1.0/1.0) box mip chains."""
    w = h = 64
    faces = [nvtt.synth.hdr_rgba16f(w, h, seed=30 + f) for f in range(6)]
    for gamma in ((2.2, 2.2), (1.0, 1.0)):
        d = nvtt.make_process_desc(nvtt.InputFormat_RGBA_16F, w, h, nvtt.Format_BC6, 1, faces=6, mip_filter=0, gamma=gamma,
                                   pixel_type=nvtt.PixelType_UnsignedFloat)
        got = ctx.process_bytes(faces, d)
        want = ref.process(faces, ref.InputFormat_RGBA_16F, w, h, ref.Format_BC6, 1, mip_filter=0, gamma=gamma,
                           pixel_type=ref.PixelType_UnsignedFloat, texture_type=ref.TextureType_Cube)
        _assert_blocks_equal(got, want, 16, "BC6H cube gamma %s" % (gamma,))


def test_bc7_level_bit_exact(nvtt, ref, ctx):
    """BC7 (AVPCL, all 8 modes): photo with alpha, opaque photo, adversarial noise / flat tiles, normal map, fp32 mip-like
    values (not multiples of 1/255), ragged sizes.  The reference asserts on values outside [0,1], so inputs stay in range."""
    rng = np.random.default_rng(23)
    s = nvtt.synth
    for (w, h) in SIZES:
        imgs = [("alpha", s.planar_from_bgra8(s.photo_bgra8(w, h, seed=1234, alpha=True))),
                ("opaque", s.planar_from_bgra8(s.photo_bgra8(w, h, seed=7))),
                ("adversarial", s.planar_from_bgra8(s.adversarial_bgra8(w, h, seed=5))),
                ("normal", s.planar_from_bgra8(s.normal_bgra8(w, h, seed=7))),
                ("float", np.clip(s.planar_from_bgra8(s.photo_bgra8(w, h, seed=5, alpha=True)) + rng.normal(0, 0.01, (4, h, w)), 0, 1).astype(np.float32))]
        for name, img in imgs:
            got = ctx.encode_level(nvtt.Format_BC7, 1, img)
            want = ref.compress_level(ref.Format_BC7, 1, img)
            _assert_blocks_equal(got, want, 16, "BC7 %s %dx%d" % (name, w, h))


def test_bc7_pipeline_bit_exact(nvtt, ref, ctx):
    """BASELINE configs[3] in small: BGRA8 texture -> BC7 with a box mip chain (and a Kaiser one), through the whole pipeline."""
    for (w, h, kw) in ((64, 64, dict(mip_filter=0)), (40, 24, dict(mip_filter=2, wrap=0))):
        img = nvtt.synth.photo_bgra8(w, h, seed=77, alpha=True)
        d = nvtt.make_process_desc(0, w, h, nvtt.Format_BC7, 1, **kw)
        got = ctx.process_bytes([img], d)
        want = ref.process([img], 0, w, h, ref.Format_BC7, 1, **kw)
        _assert_blocks_equal(got, want, 16, "BC7 pipeline %dx%d %s" % (w, h, kw))


def test_bc1a_fastest_bit_exact_where_the_reference_is_defined(nvtt, ref, ctx):
    """QuickCompress::compressDXT1a reads uninitialised stack entries for blocks that contain a texel with alpha 0
    (QuickCompressDXT.cpp:739-763: `block[num..15]` is never written but computeIndices3 walks all 16), so the reference is
    not even run-to-run stable there; every other block must match bit for bit."""
    for (w, h) in SIZES:
        for name, img in _images(nvtt, w, h):
            got = ctx.encode_level(nvtt.Format_DXT1a, 0, img).reshape(-1, 8)
            want = ref.compress_level(ref.Format_DXT1a, 0, img).reshape(-1, 8)
            a8 = (np.clip(img[3], 0, 1) * np.float32(255.0)).astype(np.int32)
            bw, bh = (w + 3) // 4, (h + 3) // 4
            defined = np.array([not (a8[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4] == 0).any() for by in range(bh) for bx in range(bw)])
            assert np.array_equal(got[defined], want[defined]), "BC1a fastest %s %dx%d" % (name, w, h)


def test_bc1a_cluster_fit_bit_exact(nvtt, ref, ctx):
    """BC1a Normal/Production (squish cluster fit in DXT1 mode): punch-through texels, all-transparent blocks, single colours
    with holes, colour weights and alpha weighting."""
    rng = np.random.default_rng(4)
    extra = []
    img = np.zeros((4, 16, 16), np.float32); img[:3] = rng.random((3, 16, 16)); extra.append(("alltransp", img))
    img = np.full((4, 16, 16), 0.5, np.float32); img[3] = (rng.random((16, 16)) > 0.5); extra.append(("singleholes", img))
    img = np.full((4, 16, 16), 0.25, np.float32); img[0, :, ::2] = 0.75; img[3] = (rng.random((16, 16)) > 0.3); extra.append(("twoholes", img))
    for (w, h) in SIZES:
        for name, img in list(_images(nvtt, w, h)) + (extra if (w, h) == SIZES[0] else []):
            for q, am, cw in ((1, 0, (1, 1, 1, 1)), (2, 1, (0.3, 0.59, 0.11, 1.0))):
                got = ctx.encode_level(nvtt.Format_DXT1a, q, img, alpha_mode=am, color_weights=cw)
                want = ref.compress_level(ref.Format_DXT1a, q, img, alpha_mode=am, color_weights=cw)
                _assert_blocks_equal(got, want, 8, "BC1a q%d %s %dx%d" % (q, name, w, h))


def test_bc3_weights_and_transparency(nvtt, ref, ctx):
    img = nvtt.synth.planar_from_bgra8(nvtt.synth.photo_bgra8(256, 256, seed=3, alpha=True))
    for cw in [(1, 1, 1, 1), (0.3, 0.59, 0.11, 1.0), (1, 0, 0, 1)]:
        for am in (nvtt.AlphaMode_None, nvtt.AlphaMode_Transparency):
            got = ctx.encode_level(nvtt.Format_BC3, 1, img, alpha_mode=am, color_weights=cw)
            want = ref.compress_level(ref.Format_BC3, 1, img, alpha_mode=am, color_weights=cw)
            _assert_blocks_equal(got, want, 16, "BC3 weights %s alphaMode %d" % (cw, am))


def test_bc1_weights_transparency_and_blacks(nvtt, ref, ctx):
    """ICBC paths that depend on the input: colour weights, alpha-weighted texels, dark blocks (3-colour + transparent
    black), out-of-range floats."""
    s = nvtt.synth
    photo = s.planar_from_bgra8(s.photo_bgra8(128, 128, seed=3, alpha=True))
    dark = np.ascontiguousarray((s.planar_from_bgra8(s.photo_bgra8(128, 128, seed=4)) * 0.2).astype(np.float32))
    rng = np.random.default_rng(9)
    oob = rng.random((4, 64, 64), dtype=np.float32) * 1.5 - 0.25
    for name, img in (("photo", photo), ("dark", dark), ("oob", oob)):
        for q in (0, 1, 2):
            for cw, am in (((1, 1, 1, 1), 0), ((0.3, 0.59, 0.11, 1.0), 1), ((1, 1, 1, 1), 1)):
                got = ctx.encode_level(nvtt.Format_BC1, q, img, alpha_mode=am, color_weights=cw)
                want = ref.compress_level(ref.Format_BC1, q, img, alpha_mode=am, color_weights=cw)
                _assert_blocks_equal(got, want, 8, "BC1 %s q%d weights %s alphaMode %d" % (name, q, cw, am))


def test_surface_ops_bit_exact(nvtt, ref, ctx):
    rng = np.random.default_rng(0)
    for (w, h) in [(64, 64), (37, 22), (33, 16), (16, 33), (7, 1), (1, 8), (2, 2), (5, 5), (256, 128)]:
        im8 = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        for wrap in (0, 1, 2):
            for filt, params in [(0, None), (1, None), (2, None), (2, (3.0, 4.0, 1.0)), (2, (2.0, 3.0, 1.5))]:
                a = ref.Surface(wrap=wrap)
                b = nvtt.Surface(ctx, wrap=wrap)
                a.set_image(0, w, h, im8)
                b.set_image(0, w, h, im8)
                assert np.array_equal(a.get().view(np.uint32), b.get().view(np.uint32)), "setImage"
                a.to_linear(2.2)
                b.to_linear(2.2)
                assert np.array_equal(a.get().view(np.uint32), b.get().view(np.uint32)), "toLinear"
                level = 0
                while True:
                    ra = a.build_next_mipmap(filt, params)
                    rb = b.build_next_mipmap(filt, params)
                    assert ra == rb
                    if not ra:
                        break
                    level += 1
                    ga, gb = a.get(), b.get()
                    assert ga.shape == gb.shape
                    assert np.array_equal(ga.view(np.uint32), gb.view(np.uint32)), \
                        "mip %d filter %d wrap %d %dx%d maxdiff %g" % (level, filt, wrap, w, h, np.abs(ga - gb).max())
                a.to_gamma(2.2)
                b.to_gamma(2.2)
                assert np.array_equal(a.get().view(np.uint32), b.get().view(np.uint32)), "toGamma"


def test_tma_mip_filter_bit_exact(nvtt, ref, ctx):
    """The TMA-fed persistent 2:1 filter kernel (k_polyphase_tma: Kaiser 13 / Mitchell 9 / Triangle 5 taps): interior tiles through
    cp.async.bulk.tensor, border tiles staged by hand for every wrap mode, partial tiles, two chained levels; and the fused
    normal-map renormalisation through the pipeline."""
    rng = np.random.default_rng(5)
    for (w, h) in [(512, 256), (328, 200), (640, 136), (1024, 1024)]:
        im = rng.random((h, w, 4), dtype=np.float32) * 1.2 - 0.1
        for wrap in (0, 1, 2):
            for kind, filt, params in [("mip", 1, None), ("mip", 2, None), ("mip", 2, (2.0, 3.0, 1.5)), ("resize", 3, None), ("resize", 2, None)]:
                if (w, h) == (1024, 1024) and (wrap != 2 or kind != "mip"):
                    continue
                a = ref.Surface(wrap=wrap)
                b = nvtt.Surface(ctx, wrap=wrap)
                a.set_image(2, w, h, im)
                b.set_image(2, w, h, im)
                for level in range(2):
                    cw, ch = w >> (level + 1), h >> (level + 1)
                    if kind == "mip":
                        assert a.build_next_mipmap(filt, params) and b.build_next_mipmap(filt, params)
                    else:
                        a.resize(cw, ch, filt)
                        b.resize(cw, ch, filt)
                    ga, gb = a.get(), b.get()
                    assert ga.shape == gb.shape == (4, ch, cw)
                    assert np.array_equal(ga.view(np.uint32), gb.view(np.uint32)), \
                        "%s filter %d %s wrap %d %dx%d level %d: %d values differ" % (kind, filt, params, wrap, w, h, level + 1, int((ga.view(np.uint32) != gb.view(np.uint32)).sum()))
    # normal map: every mip renormalised (fused into the filter kernel's epilogue)
    nm = nvtt.synth.normal_bgra8(512, 384, seed=3)
    for wrap in (0, 2):
        kw = dict(mip_filter=2, normal_map=True, wrap=wrap)
        got = ctx.process_bytes([nm], nvtt.make_process_desc(0, 512, 384, nvtt.Format_BC5, 1, **kw))
        assert np.array_equal(got, ref.process([nm], 0, 512, 384, nvtt.Format_BC5, 1, **kw))


def test_surface_misc_bit_exact(nvtt, ref, ctx):
    rng = np.random.default_rng(1)
    w, h = 96, 40
    # fp16 / fp32 / R32F inputs
    for fmt, data in [(1, rng.integers(0, 65536, (h, w, 4)).astype(np.uint16)),
                      (2, rng.normal(0, 1, (h, w, 4)).astype(np.float32)),
                      (3, rng.normal(0, 1, (h, w)).astype(np.float32))]:
        a = ref.Surface()
        b = nvtt.Surface(ctx)
        a.set_image(fmt, w, h, data)
        b.set_image(fmt, w, h, data)
        assert np.array_equal(a.get().view(np.uint32), b.get().view(np.uint32)), "setImage fmt %d" % fmt
    # RGBA16F: every one of the 2^16 half bit patterns (denormals, infinities, every NaN payload), in every channel
    allh = np.arange(65536, dtype=np.uint16).reshape(256, 256, 1)
    data = np.concatenate([allh, allh[::-1], np.roll(allh, 7, 0), allh.transpose(1, 0, 2)], 2).copy()
    a = ref.Surface()
    b = nvtt.Surface(ctx)
    a.set_image(1, 256, 256, data)
    b.set_image(1, 256, 256, data)
    assert np.array_equal(a.get().view(np.uint32), b.get().view(np.uint32)), "setImage RGBA16F: full half sweep"
    # normal-map renormalisation
    im8 = nvtt.synth.normal_bgra8(w, h)
    a = ref.Surface(normal_map=True)
    b = nvtt.Surface(ctx, normal_map=True)
    a.set_image(0, w, h, im8)
    b.set_image(0, w, h, im8)
    for op in ("expand_normals", "normalize_normal_map", "pack_normals"):
        getattr(a, op)()
        getattr(b, op)()
        assert np.array_equal(a.get().view(np.uint32), b.get().view(np.uint32)), op
    # height map -> normal map (toGreyScale + 9x9 blended Sobel toNormalMap), every wrap mode
    for wrap in (0, 1, 2):
        for (ww, hh) in ((96, 40), (7, 5), (1, 1)):
            hm = rng.integers(0, 256, (hh, ww, 4), dtype=np.uint8)
            a = ref.Surface(wrap=wrap)
            b = nvtt.Surface(ctx, wrap=wrap)
            a.set_image(0, ww, hh, hm)
            b.set_image(0, ww, hh, hm)
            a.to_grey_scale(0.3, 0.5, 0.1, 0.4)
            b.to_grey_scale(0.3, 0.5, 0.1, 0.4)
            assert np.array_equal(a.get().view(np.uint32), b.get().view(np.uint32)), "toGreyScale"
            a.to_normal_map(1.0 / 1.875, 0.5 / 1.875, 0.25 / 1.875, 0.125 / 1.875)
            b.to_normal_map(1.0 / 1.875, 0.5 / 1.875, 0.25 / 1.875, 0.125 / 1.875)
            assert np.array_equal(a.get().view(np.uint32), b.get().view(np.uint32)), "toNormalMap wrap %d %dx%d" % (wrap, ww, hh)
    # resize with every filter
    for filt in (0, 1, 2, 3):
        a = ref.Surface(wrap=1)
        b = nvtt.Surface(ctx, wrap=1)
        a.set_image(0, w, h, im8)
        b.set_image(0, w, h, im8)
        a.resize(50, 30, filt)
        b.resize(50, 30, filt)
        assert np.array_equal(a.get().view(np.uint32), b.get().view(np.uint32)), "resize filter %d" % filt
    # non-2.2 gamma goes through powf: tolerance 1e-5 relative (libm vs CUDA powf), stated by north_star
    a = ref.Surface()
    b = nvtt.Surface(ctx)
    a.set_image(0, w, h, im8)
    b.set_image(0, w, h, im8)
    a.to_linear(1.8)
    b.to_linear(1.8)
    np.testing.assert_allclose(b.get(), a.get(), rtol=1e-5, atol=1e-7)


PIPE_CASES = [
    # (name, synth, fmt, quality, kwargs)
    ("bc3_kaiser_alpha", "alpha", "BC3", 1, dict(mip_filter=2, wrap=2)),
    ("bc3_kaiser_clamp", "alpha", "BC3", 1, dict(mip_filter=2, wrap=0)),
    ("bc5_normal_kaiser", "normal", "BC5", 1, dict(mip_filter=2, wrap=2, normal_map=True)),
    ("bc5_normal_box", "normal", "BC5", 1, dict(mip_filter=0, wrap=1, normal_map=True)),
    ("bc4_box", "photo", "BC4", 1, dict(mip_filter=0)),
    ("bc1_normal_box_config0", "photo", "BC1", 1, dict(mip_filter=0, wrap=0)),
    ("bc1_fastest_box", "photo", "BC1", 0, dict(mip_filter=0)),
    ("bc1_production_box_config2", "photo", "BC1", 2, dict(mip_filter=0)),
    ("bc1_production_kaiser_transparency", "alpha", "BC1", 2, dict(mip_filter=2, alpha_mode=1)),
    ("bc3_triangle_transparency", "alpha", "BC3", 1, dict(mip_filter=1, alpha_mode=1)),
    ("bc3_box_transparency", "alpha", "BC3", 2, dict(mip_filter=0, alpha_mode=1)),
    ("bc3_no_mips_gamma1", "alpha", "BC3", 1, dict(mipmaps=False, gamma=(1.0, 1.0))),
    ("bc3_maxlevel3", "photo", "BC3", 1, dict(max_level=3)),
    ("bc5_height_to_normal_kaiser", "alpha", "BC5", 1, dict(mip_filter=2, to_normal_map=True)),
    ("bc5_production_normal_box", "normal", "BC5", 2, dict(mip_filter=0, normal_map=True)),
    ("bc3_highest_box", "alpha", "BC3", 3, dict(mip_filter=0)),
]


def _synth(nvtt, kind, w, h):
    s = nvtt.synth
    if kind == "alpha":
        return s.photo_bgra8(w, h, seed=1234, alpha=True)
    if kind == "photo":
        return s.photo_bgra8(w, h, seed=99)
    return s.normal_bgra8(w, h)


@pytest.mark.parametrize("case", PIPE_CASES, ids=[c[0] for c in PIPE_CASES])
def test_pipeline_bit_exact(nvtt, ref, ctx, case):
    name, kind, fmt_name, quality, kw = case
    fmt = getattr(nvtt, "Format_" + fmt_name)
    bs = 8 if fmt_name in ("BC4", "BC1") else 16
    for (w, h) in [(256, 256), (100, 60), (31, 17)]:
        img = _synth(nvtt, kind, w, h)
        d = nvtt.make_process_desc(nvtt.InputFormat_BGRA_8UB, w, h, fmt, quality, **kw)
        got = ctx.process_bytes([img], d)
        want = ref.process([img], 0, w, h, fmt, quality, **kw)
        _assert_blocks_equal(got, want, bs, "%s %dx%d" % (name, w, h))


def test_pipeline_faces_and_emit_order(nvtt, ref, ctx):
    w = h = 64
    faces = [nvtt.synth.photo_bgra8(w, h, seed=10 + f, alpha=True) for f in range(6)]
    d = nvtt.make_process_desc(0, w, h, nvtt.Format_BC3, 1, faces=6, mip_filter=0)
    out = ctx.process(faces, d)
    assert [(f, m) for (f, m, _, _, _) in out] == [(f, m) for f in range(6) for m in range(7)]  # face-major, mip-minor
    got = np.concatenate([b for *_, b in out])
    want = ref.process(faces, 0, w, h, ref.Format_BC3, 1, mip_filter=0, texture_type=ref.TextureType_Cube)
    assert np.array_equal(got, want)
    # face sharding used by the multi-GPU path: faces [2,4) alone give the same bytes as that slice of the whole
    d2 = nvtt.make_process_desc(0, w, h, nvtt.Format_BC3, 1, faces=6, mip_filter=0, first_face=2, last_face=4)
    part = ctx.process_bytes(faces, d2)
    per_face = got.size // 6
    assert np.array_equal(part, got[2 * per_face:4 * per_face])


def test_config2_full_size_properties(nvtt, ref, ctx):
    """BASELINE config 2 at full size (4096^2, Kaiser mips): checked through size-independent properties —
    (a) determinism, (b) every level's size, (c) block encoding is local, so random 64x64 crops of level 0 and the
    whole tail of the chain (levels <= 256^2, produced from the GPU's own fp32 mips) must match the reference."""
    w = h = 4096
    img = nvtt.synth.photo_bgra8(w, h, seed=1234, alpha=True)
    kw = dict(mip_filter=2, wrap=2)
    d = nvtt.make_process_desc(0, w, h, nvtt.Format_BC3, 1, **kw)
    out1 = ctx.process([img], d)
    out2 = ctx.process([img], d)
    assert len(out1) == 13
    for (f, m, lw, lh, b), (_, _, _, _, b2) in zip(out1, out2):
        assert (lw, lh) == (max(1, w >> m), max(1, h >> m))
        assert b.size == ((lw + 3) // 4) * ((lh + 3) // 4) * 16
        assert np.array_equal(b, b2), "non-deterministic level %d" % m
    # (c1) level-0 crops: toLinear->toGamma is per-texel, so a crop through the reference pipeline (no mips) must
    # equal the corresponding blocks of our level 0.
    lvl0 = out1[0][4].reshape(h // 4, w // 4, 16)
    rng = np.random.default_rng(0)
    for _ in range(8):
        x0 = int(rng.integers(0, w // 64)) * 64
        y0 = int(rng.integers(0, h // 64)) * 64
        crop = np.ascontiguousarray(img[y0:y0 + 64, x0:x0 + 64])
        want = ref.process([crop], 0, 64, 64, ref.Format_BC3, 1, mipmaps=False).reshape(16, 16, 16)
        assert np.array_equal(lvl0[y0 // 4:y0 // 4 + 16, x0 // 4:x0 // 4 + 16], want), "crop %d,%d" % (x0, y0)
    # (c2) the chain below 512^2: rebuild the linear fp32 level on the GPU with Surface ops, hand that exact level
    # to the reference (RGBA32F input, gamma in=1.0) and let it produce the remaining levels.
    s = nvtt.Surface(ctx, wrap=2)
    s.set_image(0, w, h, img)
    s.to_linear(2.2)
    for _ in range(3):
        assert s.build_next_mipmap(2, (3.0, 4.0, 1.0))
    lin = s.get()  # 512 x 512 linear
    il = np.ascontiguousarray(lin.transpose(1, 2, 0))
    want_tail = ref.process([il], ref.InputFormat_RGBA_32F, 512, 512, ref.Format_BC3, 1, gamma=(1.0, 2.2), **kw)
    got_tail = np.concatenate([b for (_, m, _, _, b) in out1 if m >= 3])
    assert np.array_equal(got_tail, want_tail)


def test_config0_and_config2_full_size_properties(nvtt, ref, ctx):
    """BASELINE configs[0] (BC1 Normal, 2048^2, Box mips) and configs[2] (BC1 Production, 8192^2, Box mips) at full
    size.  The 2x2 box filter is local, so any 4-aligned crop whose size is a power of two reproduces — through the
    reference's own pipeline — exactly the corresponding blocks of every level the crop still covers."""
    for (size, quality, crop) in ((2048, 1, 256), (8192, 2, 256)):
        img = nvtt.synth.photo_bgra8(size, size, seed=1234)
        if quality == 2:  # configs[2]: S1 + S5 mix
            adv = nvtt.synth.adversarial_bgra8(size // 4, size // 4, seed=5)
            img[: size // 4, : size // 4] = adv
        d = nvtt.make_process_desc(0, size, size, nvtt.Format_BC1, quality, mip_filter=0)
        out = ctx.process([img], d)
        assert len(out) == size.bit_length()
        for (f, m, lw, lh, b) in out:
            assert (lw, lh) == (max(1, size >> m), max(1, size >> m))
            assert b.size == ((lw + 3) // 4) * ((lh + 3) // 4) * 8
        rng = np.random.default_rng(1)
        spots = [(0, 0)] + [(int(rng.integers(0, size // crop)) * crop, int(rng.integers(0, size // crop)) * crop) for _ in range(5)]
        for (x0, y0) in spots:
            c = np.ascontiguousarray(img[y0:y0 + crop, x0:x0 + crop])
            want = ref.process([c], 0, crop, crop, ref.Format_BC1, quality, mip_filter=0)
            off = 0
            for m in range(crop.bit_length() - 2):  # levels where the crop is still >= 4 texels wide
                cw = crop >> m
                nb = cw // 4
                lvl = out[m][4].reshape((size >> m) // 4, (size >> m) // 4, 8)
                got = lvl[(y0 >> m) // 4:(y0 >> m) // 4 + nb, (x0 >> m) // 4:(x0 >> m) // 4 + nb]
                exp = want[off:off + nb * nb * 8].reshape(nb, nb, 8)
                assert np.array_equal(got, exp), "size %d crop (%d,%d) level %d" % (size, x0, y0, m)
                off += nb * nb * 8


def test_block_row_sharding_of_one_image(nvtt, ref, ctx):
    """BASELINE configs[2]-style tiling of ONE image over N GPUs, exercised band by band on one GPU: the bands' slices,
    put back together with the layout contract, must be byte-identical to the single-GPU chain (and to the reference)."""
    for (w, h, fmt_name, q, kw, world) in ((256, 256, "BC1", 2, dict(mip_filter=0), 8), (128, 64, "BC3", 1, dict(mip_filter=2), 2),
                                            (100, 60, "BC1", 1, dict(mip_filter=0), 4), (64, 64, "BC7", 1, dict(mip_filter=0), 4)):
        fmt = getattr(nvtt, "Format_" + fmt_name)
        img = nvtt.synth.photo_bgra8(w, h, seed=5, alpha=True)
        whole = ctx.process_bytes([img], nvtt.make_process_desc(0, w, h, fmt, q, **kw))
        parts = []
        for b in range(world):
            d = nvtt.make_process_desc(0, w, h, fmt, q, band_index=b, band_count=world, **kw)
            parts.append(ctx.process_bytes([img], d))
        d0 = nvtt.make_process_desc(0, w, h, fmt, q, band_index=0, band_count=world, **kw)
        layout = nvtt.sharding.band_layout(nvtt.lib(), d0, world)
        got = nvtt.sharding.assemble_bands(layout, parts)
        assert np.array_equal(got, whole), (w, h, fmt_name, world)
        if fmt_name != "BC7":
            assert np.array_equal(got, ref.process([img], 0, w, h, fmt, q, **kw))
        # bandOutputInPlace: every band stores its slices at their final offsets of ONE chain buffer (on several GPUs that
        # buffer is peer memory of the owner; here the same device buffer is handed to every band in turn)
        import ctypes as C
        import torch
        d_img = torch.from_numpy(img).cuda()
        nw = int(nvtt.lib().nvttb_process_whole_output_size(d0))
        assert nw == whole.size
        shared = torch.zeros(nw, dtype=torch.uint8, device="cuda")
        for b in range(world):
            d = nvtt.make_process_desc(0, w, h, fmt, q, band_index=b, band_count=world, band_output_in_place=True, **kw)
            ctx.process_to_device([d_img.data_ptr()], d, shared.data_ptr(), nw)
        ctx.synchronize()
        assert np.array_equal(shared.cpu().numpy(), whole), (w, h, fmt_name, world, "in place")
        # the big levels really are split: no band carries the whole level 0 when it divides
        if (h // 4) % world == 0 and h % 4 == 0:
            assert all(n == layout[0][0][1] for _, n, _, _ in layout[0]) and layout[0][0][1] * world == ((w + 3) // 4) * (h // 4) * (8 if fmt_name == "BC1" else 16)


def test_bc3_rgbm_bit_exact(nvtt, ref, ctx):
    """Format_BC3_RGBM (compress_dxt5_rgbm): ICBC colour block on (R,G,B)/M with weights w*M + weighted brute-force multiplier
    block; opaque, alpha-weighted, ragged sizes, HDR-ish input above 1 (saturated by the encoder)."""
    rng = np.random.default_rng(23)
    for (w, h) in SIZES:
        imgs = list(_images(nvtt, w, h))
        imgs.append(("bright", (rng.random((4, h, w)) * 1.5).astype(np.float32)))
        imgs.append(("dark", (rng.random((4, h, w)) * 0.1).astype(np.float32)))
        for name, img in imgs:
            for am in (0, 1):
                got = ctx.encode_level(12, 1, img, alpha_mode=am)
                want = ref.compress_level(12, 1, img, alpha_mode=am)
                _assert_blocks_equal(got, want, 16, "BC3_RGBM %s %dx%d alphaMode %d" % (name, w, h, am))


def _sample_blocks_strip(img, idxs, bw):
    """planar [4,h,w] -> planar [4,4,4*n]: the 4x4 tiles of the sampled blocks side by side (blocks are encoded independently)."""
    tiles = [img[:, (i // bw) * 4:(i // bw) * 4 + 4, (i % bw) * 4:(i % bw) * 4 + 4] for i in idxs]
    return np.ascontiguousarray(np.concatenate(tiles, axis=2))


def test_config3_config4_full_size_properties(nvtt, ref, ctx):
    """BASELINE configs[3] / [4] at sizes where the BC7 pipeline runs in more than one chunk (65 536 blocks) and the BC6H one over a
    whole 2048^2 face: sampled blocks (chunk boundaries included) must equal the reference's encoding of the same 4x4 tiles."""
    rng = np.random.default_rng(77)
    # BC7: 1040 x 1024 = 260 x 256 blocks = 66 560 blocks = a full chunk + a partial one (the old 32 768 boundary is sampled too)
    w, h = 1040, 1024
    img = nvtt.synth.planar_from_bgra8(nvtt.synth.photo_bgra8(w, h, seed=1234, alpha=True))
    adv = nvtt.synth.planar_from_bgra8(nvtt.synth.adversarial_bgra8(256, 256, seed=5))
    img[:, 300:556, 400:656] = adv
    got = ctx.encode_level(nvtt.Format_BC7, 1, img).reshape(-1, 16)
    bw, nb = w // 4, (w // 4) * (h // 4)
    assert got.shape[0] == nb
    idxs = [0, 32767, 32768, 32769, 65535, 65536, 65537, nb - 1] + [int(i) for i in rng.integers(0, nb, 184)]
    want = ref.compress_level(ref.Format_BC7, 1, _sample_blocks_strip(img, idxs, bw)).reshape(-1, 16)
    bad = [i for k, i in enumerate(idxs) if not np.array_equal(got[i], want[k])]
    assert not bad, "BC7 1040x1024: blocks %s differ from the reference" % bad[:8]
    # determinism across two runs of the chunked, dynamically scheduled search
    assert np.array_equal(got, ctx.encode_level(nvtt.Format_BC7, 1, img).reshape(-1, 16))
    # BC6H: one 2048^2 HDR face
    w = h = 2048
    hdr = _hdr_planar(nvtt, w, h, 11)
    got = ctx.encode_level(nvtt.Format_BC6, 1, hdr, pixel_type=nvtt.PixelType_UnsignedFloat).reshape(-1, 16)
    bw, nb = w // 4, (w // 4) * (h // 4)
    idxs = [0, nb - 1] + [int(i) for i in rng.integers(0, nb, 510)]
    want = ref.compress_level(ref.Format_BC6, 1, _sample_blocks_strip(hdr, idxs, bw), pixel_type=nvtt.PixelType_UnsignedFloat).reshape(-1, 16)
    bad = [i for k, i in enumerate(idxs) if not np.array_equal(got[i], want[k])]
    assert not bad, "BC6H 2048x2048: blocks %s differ from the reference" % bad[:8]


def test_16384_square_levels_sampled_blocks(nvtt, ref, ctx):
    """Four times the texels of the largest BASELINE config (16384^2 = 16.8 M blocks per level): index arithmetic beyond 2^31
    bytes of planar input.  Sampled blocks of BC1 Fastest / Normal, BC3 and BC5 must equal the reference's encoding of the same tiles."""
    rng = np.random.default_rng(5)
    base = nvtt.synth.planar_from_bgra8(nvtt.synth.photo_bgra8(2048, 2048, seed=99, alpha=True))
    n = 16384
    img = np.empty((4, n, n), np.float32)
    for ty in range(8):
        for tx in range(8):
            # every tile differs: rolled copy plus a small per-tile offset, kept inside [0, 1]
            t = np.roll(base, (17 * ty + 3, 29 * tx + 5), axis=(1, 2))
            img[:, ty * 2048:(ty + 1) * 2048, tx * 2048:(tx + 1) * 2048] = np.clip(t * 0.9 + 0.0125 * ((ty * 8 + tx) % 8), 0.0, 1.0)
    bw = n // 4
    nb = bw * bw
    idxs = [0, bw - 1, nb - bw, nb - 1, nb // 2 + 12345] + [int(i) for i in rng.integers(0, nb, 123)]
    strip = _sample_blocks_strip(img, idxs, bw)
    for fmt, q in ((nvtt.Format_BC1, 0), (nvtt.Format_BC1, 1), (nvtt.Format_BC3, 1), (nvtt.Format_BC5, 1)):
        bs = 8 if fmt == nvtt.Format_BC1 else 16
        got = ctx.encode_level(fmt, q, img).reshape(-1, bs)
        assert got.shape[0] == nb
        want = ref.compress_level(fmt, q, strip).reshape(-1, bs)
        bad = [i for k, i in enumerate(idxs) if not np.array_equal(got[i], want[k])]
        assert not bad, "16384^2 fmt %d q%d: blocks %s differ from the reference" % (fmt, q, bad[:8])


def test_gamma_dense_float_sweep(nvtt, ref, ctx):
    """toLinear / toGamma (powf_11_5 / powf_5_11 table x polynomial, Gamma.cpp:311-354) on a dense sweep of the fp32 bit
    patterns of [0, 1]: every 64th pattern per channel with a different residue in R, G, B (50 M values), plus values just
    outside the range, negatives, infinities and NaNs; bit-exact against the reference."""
    n = 4096
    base = np.arange(n * n, dtype=np.uint64) * 64
    top = np.uint64(0x3F800000)
    img = np.empty((n, n, 4), np.uint32)
    for c, off in enumerate((0, 21, 42)):
        img[..., c] = np.minimum(base + off, top).astype(np.uint32).reshape(n, n)
    special = np.array([1.0000001, 1.5, 2.0, 255.0, 65504.0, 3.4e38, -0.0, -1e-30, -0.5, -2.0, np.inf, -np.inf, np.nan, 1e-45, 1.1754942e-38],
                       np.float32).view(np.uint32)
    img[..., 3] = 0x3F800000
    img[0, :special.size, 0] = special
    img[1, :special.size, 1] = special
    img[2, :special.size, 2] = special
    data = img.view(np.float32)
    for op in ("to_linear", "to_gamma"):
        a = ref.Surface()
        b = nvtt.Surface(ctx)
        a.set_image(2, n, n, data)
        b.set_image(2, n, n, data)
        getattr(a, op)(2.2)
        getattr(b, op)(2.2)
        ga, gb = a.get().view(np.uint32), b.get().view(np.uint32)
        assert np.array_equal(ga, gb), "%s: %d of %d values differ" % (op, int((ga != gb).sum()), ga.size)
        del a, b, ga, gb


def test_half_from_float_dense_sweep(nvtt, ref, ctx):
    """nv::half_from_float (Half.cpp:378-441) through the RGBA16F pixel-format writer on every 257th fp32 bit pattern of the whole
    32-bit range (16.7 M values: all exponents, denormals, both signs, infinities, NaN payloads), byte-identical to the reference."""
    w, h = 4096, 1024
    bits = (np.arange(w * h * 4, dtype=np.uint64) * 257 % (1 << 32)).astype(np.uint32).reshape(h, w, 4)
    vals = bits.view(np.float32)
    planar = np.ascontiguousarray(np.moveaxis(vals, 2, 0))
    got = ctx.convert_level(planar, sizes=(16, 16, 16, 16), pixel_type=4)
    want = ref.process([vals], 2, w, h, 0, 1, mipmaps=False, gamma=(1.0, 1.0), pixel_sizes=(16, 16, 16, 16), pixel_type=4)
    assert got.size == want.size
    assert np.array_equal(got, want), "%d of %d halfs differ" % (int((got.view(np.uint16) != want.view(np.uint16)).sum()), got.size // 2)
