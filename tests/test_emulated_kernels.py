"""CPU tests of the product's device code under the lock-step SIMT emulator (tests/simt_emu, a debug aid — NOT a
fallback and NOT the oracle): the same .cuh kernels that nvcc compiles for sm_100a are compiled with g++ and checked
against the golden vectors.  Catches kernel-logic regressions in this GPU-less container before a gpurun call."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_cases as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "simt_emu", "_build", "libnvb_emu.so")


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(SO):
        pytest.skip("emulator build missing (run __graft_entry__.build())")
    return C.CDLL(SO)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


def _encode(E, fmt, q, img, am, cw, pt=0):
    _, h, w = img.shape
    nb = ((w + 3) // 4) * ((h + 3) // 4)
    p = C.c_void_p(img.ctypes.data)
    if fmt == 1:
        out = np.zeros(nb * 8, np.uint8)
        c = (C.c_float * 3)(*cw[:3])
        E.emu_bc1(p, w, h, c, {0: 1, 1: 8, 2: 9, 3: 8}[q], int(am == 1), C.c_void_p(out.ctypes.data), 0)
    elif fmt == 10:
        out = np.zeros(nb * 16, np.uint8)
        E.emu_bc6(p, w, h, int(pt not in (0, 2, 5)), int(am == 1), C.c_void_p(out.ctypes.data), 0)
    elif fmt == 11:
        out = np.zeros(nb * 16, np.uint8)
        E.emu_bc7(p, w, h, C.c_void_p(out.ctypes.data), 0, 255, None, None)
    elif fmt == 6:
        out = np.zeros(nb * 8, np.uint8)
        E.emu_alpha_blocks(p, w, h, 0, C.c_void_p(out.ctypes.data), 8, 0, 0, int(q >= 2))
    elif fmt == 7:
        out = np.zeros(nb * 16, np.uint8)
        E.emu_alpha_blocks(p, w, h, 0, C.c_void_p(out.ctypes.data), 16, 0, 0, int(q >= 2))
        E.emu_alpha_blocks(p, w, h, 1, C.c_void_p(out.ctypes.data), 16, 8, 0, int(q >= 2))
    elif fmt == 2:
        out = np.zeros(nb * 8, np.uint8)
        E.emu_bc3_color_ex(p, w, h, (C.c_float * 3)(*cw[:3]), int(am == 1), C.c_void_p(out.ctypes.data), 8, 0, 0, 2)
    elif fmt in (3, 4, 5) and q == 0:
        out = np.zeros(nb * 16, np.uint8)
        o = C.c_void_p(out.ctypes.data)
        if fmt == 3:
            E.emu_alpha_blocks(p, w, h, 3, o, 16, 0, 0, 2)
        else:
            E.emu_alpha_blocks(p, w, h, 3 if fmt == 4 else 0, o, 16, 0, 0, 0)
        E.emu_dxt1_quick(p, w, h, 0, int(fmt == 5), o, 16, 8, 0)
    elif fmt == 3:
        out = np.zeros(nb * 16, np.uint8)
        E.emu_alpha_blocks(p, w, h, 3, C.c_void_p(out.ctypes.data), 16, 0, 0, 2)
        E.emu_bc3_color_ex(p, w, h, (C.c_float * 3)(*cw[:3]), int(am == 1), C.c_void_p(out.ctypes.data), 16, 8, 0, 0)
    elif fmt == 5:
        out = np.zeros(nb * 16, np.uint8)
        E.emu_alpha_blocks(p, w, h, 0, C.c_void_p(out.ctypes.data), 16, 0, 0, 0)
        E.emu_bc3_color_ex(p, w, h, (C.c_float * 3)(0, 1, 0), int(am == 1), C.c_void_p(out.ctypes.data), 16, 8, 0, 1)
    else:
        out = np.zeros(nb * 16, np.uint8)
        E.emu_alpha_blocks(p, w, h, 3, C.c_void_p(out.ctypes.data), 16, 0, 0, int(q == 3))
        m = (C.c_float * 3)(*cw[:3])
        E.emu_bc3_color(p, w, h, m, int(am == 1), C.c_void_p(out.ctypes.data), 16, 8, 0)
    return out


def test_emulated_encoders_match_golden(emu, golden):
    for key, (kind, w, h, fmt, q, am, cw, pt) in G.level_cases().items():
        if (w, h) != (13, 7) and kind != "photo" and fmt not in (10, 11):
            continue  # keep the CPU suite short: ragged size for every input kind, full size for one
        img = G.make_input(kind, w, h, planar=True)
        got = _encode(emu, fmt, q, img, am, cw, pt)
        assert np.array_equal(got, golden[key]), key


def _run_child(env, code):
    import subprocess
    import sys
    e = dict(os.environ)
    e.update(env)
    e["PYTHONPATH"] = ROOT + os.pathsep + os.path.join(ROOT, "tests") + os.pathsep + e.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


_CHILD = """
import ctypes as C, hashlib, numpy as np
import golden_cases as G
E = C.CDLL(%r)
img = G.make_input(%r, %d, %d, planar=True)
nb = ((%d + 3) // 4) * ((%d + 3) // 4)
out = np.zeros(nb * 16, np.uint8)
if %d == 11:
    errs = np.zeros(8 * nb, np.float32); cands = np.zeros(8 * nb * 16, np.uint8)
    E.emu_bc7(C.c_void_p(img.ctypes.data), %d, %d, C.c_void_p(out.ctypes.data), 0, 255, C.c_void_p(errs.ctypes.data), C.c_void_p(cands.ctypes.data))
    print(hashlib.sha1(out.tobytes() + errs.tobytes() + cands.tobytes()).hexdigest())
else:
    E.emu_bc6(C.c_void_p(img.ctypes.data), %d, %d, 0, 0, C.c_void_p(out.ctypes.data), 0)
    print(hashlib.sha1(out.tobytes()).hexdigest())
"""


def test_search_designs_agree(emu):
    """The product's searcher state machines against the two older mappings of the same search (thread per candidate, warp
    per candidate for BC7; thread per (block, kind) for BC6H): blocks, per-mode errors and per-mode candidates identical."""
    def child(kind, w, h, fmt):
        return _CHILD % (SO, kind, w, h, w, h, fmt, w, h, w, h)
    for kind, w, h in (("photo", 20, 12), ("adv", 13, 7)):
        ref = _run_child({"NVB_EMU_BC7": "scalar"}, child(kind, w, h, 11))
        assert _run_child({"NVB_EMU_BC7": ""}, child(kind, w, h, 11)) == ref, (kind, "search")
    assert _run_child({"NVB_EMU_BC7": "coop"}, child("photo", 8, 8, 11)) == _run_child({"NVB_EMU_BC7": "scalar"}, child("photo", 8, 8, 11))
    for kind, w, h in (("hdr", 32, 24), ("photo", 13, 7)):
        assert _run_child({"NVB_EMU_BC6": "scalar"}, child(kind, w, h, 10)) == _run_child({"NVB_EMU_BC6": ""}, child(kind, w, h, 10)), kind


def _pixel_layout(w, kw):
    """capi.cu: pixel_layout() restated for the emulator test: (kind, bitCount, sizes, shifts, pitch, aligned)."""
    pt, align = kw.get("pixel_type", 0), kw.get("pitch_alignment", 1)
    size, shift = [0, 0, 0, 0], [0, 0, 0, 0]
    if kw.get("sizes") is not None:
        size = list(kw["sizes"])
        bits = sum(size)
        shift = [size[1] + size[2] + size[3], size[2] + size[3], size[3], 0]
    else:
        bits, *masks = kw.get("masks") or (32, 0xFF0000, 0xFF00, 0xFF, 0xFF000000)
        for i, m in enumerate(masks):
            if m:
                while not (m >> shift[i]) & 1:
                    shift[i] += 1
                while (m >> (shift[i] + size[i])) & 1:
                    size[i] += 1
    if pt == 4:
        kind, aligned = 2, all(s in (0, 16, 32) for s in size)
        shift = [0, 0, 0, 0]
    else:
        kind, aligned = {0: 0, 2: 1}.get(pt, 3), bits % 8 == 0
    abits = 8 * align
    pitch = ((w * bits + abits - 1) // abits * abits + 7) // 8
    return kind, bits, size, shift, pitch, aligned


def test_emulated_pixel_format_kernels_match_golden(emu, nvtt):
    """The three Format_RGBA kernels under the CPU emulator against golden_v3.npz (reference output)."""
    g3 = np.load(os.path.join(ROOT, "tests", "golden", "golden_v3.npz"))
    s = nvtt.synth
    U4 = C.c_uint * 4
    emu.emu_pixel_format.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_int, U4, U4, C.c_int, C.c_int]
    ran = set()
    for key, (w, h, kw) in G.pixel_format_cases().items():
        img = np.ascontiguousarray(s.planar_from_bgra8(s.photo_bgra8(w, h, seed=w * 3 + h, alpha=True)), dtype=np.float32)
        kind, bits, size, shift, pitch, aligned = _pixel_layout(w, kw)
        nbytes = bits // 8
        mode = 0
        if aligned and kind <= 1 and nbytes in (1, 2, 4):
            mode = nbytes
        elif kind == 2 and size == [16] * 4:
            mode = 8
        elif kind == 2 and size == [32] * 4:
            mode = 16
        paths = [2] + ([1] if aligned else []) + ([0] if mode and w % 4 == 0 and pitch == w * nbytes else [])
        for path in paths:  # the per-scanline stream is the general kernel: it must agree everywhere
            out = np.full(pitch * h, 0xCD, np.uint8)
            emu.emu_pixel_format(img.ctypes.data, w, h, out.ctypes.data, pitch, bits, kind, U4(*size), U4(*shift), path, mode)
            assert out.size == g3[key].size and np.array_equal(out, g3[key]), (key, path)
            ran.add(path)
    assert ran == {0, 1, 2}


def test_emulated_rgb9e5_matches_reference(emu, ref):
    """k_pixel_format / k_pixel_format_rows with the R9G9B9E5 writer (kind 4) under the CPU emulator against the reference."""
    from test_oracle import rgb9e5_probe_values
    vals, w, h = rgb9e5_probe_values()
    planar = np.ascontiguousarray(np.moveaxis(vals, 2, 0))
    want = ref.process([vals], 2, w, h, 0, 1, mipmaps=False, gamma=(1.0, 1.0), pixel_sizes=(9, 9, 9, 5), pixel_type=6)
    U4 = C.c_uint * 4
    emu.emu_pixel_format.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_int, U4, U4, C.c_int, C.c_int]
    for path in (1, 2):
        out = np.full(4 * w * h, 0xCD, np.uint8)
        emu.emu_pixel_format(planar.ctypes.data, w, h, out.ctypes.data, 4 * w, 32, 4, U4(9, 9, 9, 5), U4(23, 14, 5, 0), path, 0)
        assert np.array_equal(out, want), (path, int(np.flatnonzero(out != want)[0]) // 4)


def test_emulated_set_image_x4_matches_scalar_kernel(emu):
    """k_set_image_bgra8_x4 (256-entry tables, four pixels per thread) against k_set_image, with and without the fused toLinear."""
    rng = np.random.default_rng(9)
    n = 4096 * 3
    src = rng.integers(0, 256, (n, 4), dtype=np.uint8)
    src[:256, :] = np.arange(256, dtype=np.uint8)[:, None]  # every byte value in every channel
    for lin in (0, 1):
        a = np.zeros(4 * n, np.float32)
        b = np.zeros(4 * n, np.float32)
        emu.emu_set_image(C.c_void_p(src.ctypes.data), C.c_void_p(a.ctypes.data), n, 0, lin)
        emu.emu_set_image_x4(C.c_void_p(src.ctypes.data), C.c_void_p(b.ctypes.data), n, lin)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), lin
