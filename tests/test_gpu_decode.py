"""GPU BCn decoders (Surface::setImage2D) and error metrics (nvtt::rmsError / rmsAlphaError) against the reference:
decoded texels must be bit-identical for blocks produced by the encoders AND for arbitrary block bytes (every BC7 /
BC6H mode, reserved modes, 3-colour DXT1 blocks inside BC2/BC3); the metrics agree to 1e-6 relative (the reference adds
fp32 terms into one double sequentially, the GPU reduces the same terms in double in a different order)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _harness(path):
    L = C.CDLL(path)
    L.ref_decode_ex.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_rms_error.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.ref_angular_error.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    L.ref_cielab_error.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    return L


@pytest.fixture(scope="module")
def libs(ref):
    ours = os.path.join(ROOT, "tests", "_build", "libnvtt_b200_harness.so")
    theirs = os.path.join(ROOT, "oracle", "_ref", "libnvtt_ref.so")
    if not os.path.exists(ours) or not os.path.exists(theirs):
        pytest.fail("harness libraries missing: run __graft_entry__.build()")
    return _harness(ours), _harness(theirs)


def _decode(L, fmt, decoder, w, h, blocks):
    out = np.zeros((4, h, w), np.float32)
    blocks = np.ascontiguousarray(blocks)
    assert L.ref_decode_ex(fmt, decoder, w, h, blocks.ctypes.data, out.ctypes.data) == 1
    return out


FORMATS = {"BC1": (1, 8), "BC2": (3, 16), "BC3": (4, 16), "BC3n": (5, 16), "BC4": (6, 8), "BC5": (7, 16), "BC6": (10, 16), "BC7": (11, 16)}


def test_decode_encoded_levels_bit_exact(nvtt, ref, ctx, libs):
    ours, theirs = libs
    s = nvtt.synth
    img = s.planar_from_bgra8(s.photo_bgra8(52, 36, seed=3, alpha=True))
    hdr = np.ascontiguousarray(np.moveaxis(s.hdr_rgba16f(52, 36, seed=5).astype(np.float32), 2, 0))
    for name, (fmt, _) in FORMATS.items():
        src = hdr if name == "BC6" else img
        kw = dict(pixel_type=5) if name == "BC6" else {}
        blocks = ctx.encode_level(fmt, 1, src, **kw)
        for decoder in ((0,) if name in ("BC6", "BC7") else (0, 1, 2)):
            a = _decode(ours, fmt, decoder, 52, 36, blocks)
            b = _decode(theirs, fmt, decoder, 52, 36, blocks)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, decoder)
    # the C-ABI view gives the same texels as the C++ mirror
    blocks = ctx.encode_level(11, 1, img)
    surf = nvtt.Surface(ctx)
    surf.set_image_2d(11, 52, 36, blocks)
    assert np.array_equal(surf.get().view(np.uint32), _decode(theirs, 11, 0, 52, 36, blocks).view(np.uint32))


def test_decode_arbitrary_blocks_bit_exact(libs):
    ours, theirs = libs
    rng = np.random.default_rng(99)
    w, h = 64, 60
    nb = (w // 4) * (h // 4)
    for name, (fmt, bs) in FORMATS.items():
        blocks = rng.integers(0, 256, nb * bs, dtype=np.uint8)
        if name == "BC7":
            blocks = blocks.reshape(nb, 16)
            blocks[:8, 0] = 0                       # reserved mode: all-zero texels
            # mode 0 is left out here: the reference cannot decode it (its read_header never consumes the mode bit and the
            # assert that follows ends the process, src/bc7/avpcl_mode0.cpp:246-250); see test_decode_bc7_mode0_by_the_format
            blocks[:, 0] &= np.uint8(0xFE)
            blocks[blocks[:, 0] == 0, 0] = 2
            for m in range(1, 8):                   # every other mode, several times
                blocks[8 + m::16, 0] = (blocks[8 + m::16, 0] & np.uint8(0xFF ^ ((1 << (m + 1)) - 1))) | np.uint8(1 << m)
            blocks = blocks.reshape(-1)
        if name == "BC6":
            blocks = blocks.reshape(nb, 16)
            modes = [0x00, 0x01, 0x02, 0x06, 0x0a, 0x0e, 0x12, 0x16, 0x1a, 0x1e, 0x03, 0x07, 0x0b, 0x0f, 0x13, 0x17, 0x1b, 0x1f]
            for i, m in enumerate(modes):
                mask = 0x03 if m < 2 else 0x1F
                blocks[i::len(modes), 0] = (blocks[i::len(modes), 0] & np.uint8(0xFF ^ mask)) | np.uint8(m)
            blocks = blocks.reshape(-1)
        for decoder in ((0,) if name in ("BC6", "BC7") else (0, 1, 2)):
            a = _decode(ours, fmt, decoder, w, h, blocks)
            b = _decode(theirs, fmt, decoder, w, h, blocks)
            bad = int((a.view(np.uint32) != b.view(np.uint32)).sum())
            assert bad == 0, (name, decoder, bad)


def test_decode_ragged_size(libs):
    ours, theirs = libs
    rng = np.random.default_rng(5)
    for (w, h) in ((13, 7), (1, 1), (5, 9)):
        nb = ((w + 3) // 4) * ((h + 3) // 4)
        for name in ("BC1", "BC3", "BC7"):
            fmt, bs = FORMATS[name]
            blocks = rng.integers(0, 256, nb * bs, dtype=np.uint8)
            if name == "BC7":
                blocks[0::16] = (blocks[0::16] & np.uint8(0xFC)) | np.uint8(2)  # mode 1
            a = _decode(ours, fmt, 0, w, h, blocks)
            b = _decode(theirs, fmt, 0, w, h, blocks)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, w, h)


def _table(name, text):
    import re
    body = re.search(name + r"\[\d+\]\s*=\s*\{(.*?)\};", text, re.S).group(1)
    return [int(t.rstrip("u"), 0) for t in re.findall(r"0x[0-9a-fA-F]+u?|\d+", body)]


def test_decode_bc7_mode0_by_the_format(nvtt, ctx):
    """BC7 mode 0 (3 subsets, 4-bit endpoints + unique p-bits, 3-bit indices) decoded by a plain restatement of the format."""
    text = open(os.path.join(ROOT, "nvidia-texture-tools_b200", "csrc", "kernels", "bc67_tables.cuh")).read()
    shape3, anchor3 = _table("kShape3", text), _table("kAnchor3", text)
    w7 = [0, 9, 18, 27, 37, 46, 55, 64]
    rng = np.random.default_rng(7)
    nb = 256
    blocks = rng.integers(0, 256, (nb, 16), dtype=np.uint8)
    blocks[:, 0] |= 1
    want = np.zeros((nb, 16, 4), np.float32)
    for n in range(nb):
        bits = int.from_bytes(bytes(blocks[n]), "little")
        pos = [1]

        def rd(k):
            v = (bits >> pos[0]) & ((1 << k) - 1)
            pos[0] += k
            return v
        shape = rd(4)
        ep = np.zeros((3, 2, 3), int)
        for ch in range(3):
            for r in range(3):
                ep[r, 0, ch] = rd(4)
                ep[r, 1, ch] = rd(4)
        lsb = [[rd(1), rd(1)] for _ in range(3)]
        anchors = {0, anchor3[2 * shape], anchor3[2 * shape + 1]}
        for t in range(16):
            idx = rd(2 if t in anchors else 3)
            r = (shape3[shape] >> (2 * t)) & 3
            for ch in range(3):
                a = (ep[r, 0, ch] << 1) | lsb[r][0]
                b = (ep[r, 1, ch] << 1) | lsb[r][1]
                a = (a << 3) | (a >> 2)
                b = (b << 3) | (b >> 2)
                want[n, t, ch] = np.float32((a * w7[7 - idx] + b * w7[idx] + 32) >> 6) / np.float32(255.0)
            want[n, t, 3] = 1.0
    surf = nvtt.Surface(ctx)
    surf.set_image_2d(11, 64, 64, blocks.reshape(-1))
    got = surf.get()  # [4, 64, 64]
    got = got.reshape(4, 16, 4, 16, 4).transpose(1, 3, 2, 4, 0).reshape(nb, 16, 4)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_rms_error_matches_reference(nvtt, ctx, libs):
    ours, theirs = libs
    s = nvtt.synth
    w, h = 256, 128
    bgra = s.photo_bgra8(w, h, seed=11, alpha=True)
    planar = s.planar_from_bgra8(bgra)
    rgba = np.ascontiguousarray(np.moveaxis(planar, 0, 2))  # RGBA32F interleaved
    for fmt in (1, 4, 7):  # not BC7: the reference aborts on the first mode-0 block it has to decode
        blocks = ctx.encode_level(fmt, 1, planar)
        for am in (0, 1):
            vals = []
            for L in (ours, theirs):
                a, b = C.c_float(), C.c_float()
                assert L.ref_rms_error(fmt, w, h, blocks.ctypes.data, rgba.ctypes.data, am, C.byref(a), C.byref(b)) == 1
                vals.append((a.value, b.value))
            (a0, b0), (a1, b1) = vals
            assert a1 > 0 and abs(a0 - a1) <= 1e-6 * a1, (fmt, am, a0, a1)
            assert abs(b0 - b1) <= 1e-6 * max(b1, 1e-30), (fmt, am, b0, b1)


def test_angular_error_matches_reference(nvtt, ctx, libs):
    """nvtt::angularError of a BC5 / BC3n encoded normal map: acosf differs between glibc and CUDA, hence 1e-5 relative."""
    ours, theirs = libs
    w, h = 128, 96
    planar = nvtt.synth.planar_from_bgra8(nvtt.synth.normal_bgra8(w, h, seed=3))
    rgba = np.ascontiguousarray(np.moveaxis(planar, 0, 2))
    for fmt in (7, 1):
        blocks = ctx.encode_level(fmt, 1, planar)
        vals = []
        for L in (ours, theirs):
            v = C.c_float()
            assert L.ref_angular_error(fmt, w, h, blocks.ctypes.data, rgba.ctypes.data, C.byref(v)) == 1
            vals.append(v.value)
        assert vals[1] > 0 and abs(vals[0] - vals[1]) <= 1e-5 * vals[1], (fmt, vals)


def test_cielab_error_matches_reference(nvtt, ctx, libs):
    """nvtt::cieLabError of BC1 / BC3 encoded photos: powf differs between glibc and CUDA, hence 1e-4 relative."""
    ours, theirs = libs
    w, h = 128, 96
    planar = nvtt.synth.planar_from_bgra8(nvtt.synth.photo_bgra8(w, h, seed=13, alpha=True))
    rgba = np.ascontiguousarray(np.moveaxis(planar, 0, 2))
    for fmt in (1, 4):
        blocks = ctx.encode_level(fmt, 1, planar)
        vals = []
        for L in (ours, theirs):
            v = C.c_float()
            assert L.ref_cielab_error(fmt, w, h, blocks.ctypes.data, rgba.ctypes.data, C.byref(v)) == 1
            vals.append(v.value)
        assert vals[1] > 0 and abs(vals[0] - vals[1]) <= 1e-4 * vals[1], (fmt, vals)
