"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/nvtt_b200.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_and_library_agree(nvtt):
    hdr = open(os.path.join(ROOT, "include", "nvtt_b200.h")).read()
    declared = sorted(set(re.findall(r"NVTTB_API[^;(]*?\b(nvttb_\w+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = nvtt.lib()
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, "declared in the header but not exported: %s" % missing
    assert sorted(nvtt.EXPORTS) == declared


def test_build_variants(nvtt):
    """The strict (bit-exact) and the opt-in fast-math build export the same C ABI and say which one they are."""
    import ctypes as C
    L = nvtt.lib()
    assert L.nvttb_build_variant() == b"strict"
    fast = os.path.join(os.path.dirname(nvtt.capi.LIB_PATH), "libnvtt_b200_fastmath.so")
    assert os.path.exists(fast), "run __graft_entry__.build()"
    F = C.CDLL(fast)
    F.nvttb_build_variant.restype = C.c_char_p
    assert F.nvttb_build_variant() == b"fastmath"
    missing = [n for n in nvtt.EXPORTS if not hasattr(F, n)]
    assert not missing, missing


def test_sizes_and_mip_counts(nvtt):
    L = nvtt.lib()
    assert L.nvttb_level_size(nvtt.Format_BC1, 2048, 2048) == 512 * 512 * 8
    assert L.nvttb_level_size(nvtt.Format_BC3, 5, 3) == 2 * 1 * 16
    assert L.nvttb_level_size(nvtt.Format_BC4, 1, 1) == 8
    d = nvtt.make_process_desc(nvtt.InputFormat_BGRA_8UB, 2048, 2048, nvtt.Format_BC3, nvtt.Quality_Normal)
    assert L.nvttb_process_mip_count(d) == 12
    d = nvtt.make_process_desc(nvtt.InputFormat_BGRA_8UB, 37, 22, nvtt.Format_BC3, nvtt.Quality_Normal)
    assert L.nvttb_process_mip_count(d) == 6  # 37x22,18x11,9x5,4x2,2x1,1x1
    sizes = [(37, 22), (18, 11), (9, 5), (4, 2), (2, 1), (1, 1)]
    assert L.nvttb_process_output_size(d) == sum(((w + 3) // 4) * ((h + 3) // 4) * 16 for w, h in sizes)
    d = nvtt.make_process_desc(nvtt.InputFormat_BGRA_8UB, 256, 256, nvtt.Format_BC3, nvtt.Quality_Normal, max_level=3)
    assert L.nvttb_process_mip_count(d) == 3
    d = nvtt.make_process_desc(nvtt.InputFormat_BGRA_8UB, 256, 256, nvtt.Format_BC3, nvtt.Quality_Normal, mipmaps=False)
    assert L.nvttb_process_mip_count(d) == 1


def test_no_cpu_fallback(nvtt):
    """Without a device the context cannot be created; nothing in the product computes on the CPU."""
    if nvtt.lib().nvttb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nvtt.NvttbError) as e:
        nvtt.Context(0)
    assert e.value.code == 4  # Error_CudaError


def test_host_library_exports_nvtt_api():
    """The C++ drop-in (lib/libnvtt.so) exports the nvtt:: classes the reference's callers link against."""
    import subprocess
    lib = os.path.join(ROOT, "nvidia-texture-tools_b200", "lib", "libnvtt.so")
    if not os.path.exists(lib):
        pytest.skip("host library not built")
    syms = subprocess.run(["nm", "-DC", "--defined-only", lib], capture_output=True, text=True).stdout
    for want in ("nvtt::Compressor::process(", "nvtt::Compressor::compress(int, int, int, int, int, float const*",
                 "nvtt::Compressor::outputHeader(", "nvtt::Compressor::estimateSize(", "nvtt::InputOptions::setMipmapData(",
                 "nvtt::CompressionOptions::setFormat(", "nvtt::OutputOptions::setOutputHandler(", "nvtt::Surface::buildNextMipmap(",
                 "nvtt::Surface::toGamma(", "nvtt::version()"):
        assert want in syms, want


def test_dds_header_matches_reference(ref):
    """The DDS / DDS10 / KTX header written by our outputHeader (host code, no GPU needed) is byte-identical to the reference's."""
    import ctypes as C
    import numpy as np
    so = os.path.join(ROOT, "tests", "_build", "libnvtt_b200_harness.so")
    if not os.path.exists(so):
        pytest.skip("harness against the drop-in header not built")
    ours = C.CDLL(so)
    ours.ref_process.restype = C.c_long
    ours.ref_process.argtypes = [C.POINTER(ref.RefProcessDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_long]
    img = np.zeros((8, 8, 4), np.uint8)
    for fmt in (ref.Format_BC1, ref.Format_DXT1a, ref.Format_BC2, ref.Format_BC3, ref.Format_BC3n, ref.Format_BC4, ref.Format_BC5,
                ref.Format_BC6, ref.Format_BC7, ref.Format_BC3_RGBM):
        for container in (ref.Container_DDS, ref.Container_DDS10, ref.Container_KTX):
            for normal in (False, True):
                for (ttype, faces) in ((ref.TextureType_2D, 1), (ref.TextureType_Cube, 6)):
                    for mips in (True, False):
                        if container == ref.Container_DDS and fmt in (ref.Format_BC6, ref.Format_BC7):
                            continue  # reference reports Error_UnsupportedOutputFormat
                        try:
                            want = ref.process([img] * faces, 0, 8, 8, fmt, 1, header=True, container=container, normal_map=normal,
                                               texture_type=ttype, mipmaps=mips, pixel_type=5 if fmt == ref.Format_BC6 else 0)
                        except RuntimeError:
                            continue
                        hs = 148 if container == ref.Container_DDS10 else 64 if container == ref.Container_KTX else 128
                        # ours: without a GPU process() fails after the header was written; capture the header through
                        # a direct outputHeader call instead
                        got = _our_header(ours, fmt, container, normal, ttype, faces, 4 if mips else 1)
                        assert got is not None
                        assert bytes(want[:hs]) == got, (fmt, container, normal, ttype, mips)


def _our_header(ours, fmt, container, normal, ttype, faces, mipcount):
    import ctypes as C
    if not hasattr(ours, "ref_output_header"):
        return None
    buf = (C.c_ubyte * 256)()
    ours.ref_output_header.restype = C.c_long
    n = ours.ref_output_header(ttype, 8, 8, faces if ttype == 3 else 1, mipcount, int(normal), fmt, container, buf, 256)
    return bytes(buf[:n]) if n > 0 else None


def test_div7_identity():
    """alpha_div7 (bc_alpha.cuh): k * RN(1/7) corrected by two fused multiply-adds equals the IEEE quotient k / 7.0f for
    k = 1..6 - the only operands the DXT5 alpha refit ever divides (QuickCompressDXT.cpp optimizeAlpha8 weights)."""
    import numpy as np
    f32 = np.float32
    y = f32(1.0) / f32(7.0)
    for k in range(1, 7):
        kk = f32(k)
        q = f32(kk * y)
        r = f32(np.float64(kk) - np.float64(7.0) * np.float64(q))      # fma(-7, q, k): exact
        q2 = f32(np.float64(q) + np.float64(r) * np.float64(y))       # fma(r, y, q): |terms| small, double is exact enough
        assert q2 == f32(kk / f32(7.0)), k


def test_div255_identity():
    """icbc_u8_to_float (bc1_icbc.cuh): v * RN(1/255) corrected by two fused multiply-adds equals the IEEE quotient
    float(v) / 255.0f for every v in 0..255 - the only operands the BC1 palette ever divides (icbc.h evaluate_palette)."""
    import numpy as np
    from fractions import Fraction
    f32 = np.float32
    r = np.frombuffer(np.uint32(0x3b808081).tobytes(), f32)[0]
    assert r == f32(1.0) / f32(255.0)

    def rn32(fr):  # correctly rounded float32 of an exact rational (no ties occur here; nearest of the three neighbours)
        c = f32(float(fr))
        return min((c, np.nextafter(c, f32(np.inf)), np.nextafter(c, f32(-np.inf))), key=lambda z: abs(Fraction(float(z)) - fr))

    for v in range(256):
        x = f32(v)
        q = f32(x * r)
        rem = rn32(Fraction(float(q)) * -255 + Fraction(float(x)))
        got = rn32(Fraction(float(r)) * Fraction(float(rem)) + Fraction(float(q)))
        assert got == f32(x / f32(255.0)), v
