"""The DDS / DDS10 container reader (nvttb_dds_describe / nvttb_dds_surface: host code, no GPU needed to parse): files written by
the REFERENCE library (its own header writer) are located surface by surface; and - on the GPU - a pre-made mip chain read
this way feeds InputOptions::setMipmapData for levels > 0 with the same result as the reference fed the same levels."""
import numpy as np
import pytest


def _ref_or_skip():
    import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref not built")
    return refapi


def _level_sizes(w, h, bytes_per_block=None, bpp=None):
    out = []
    while True:
        out.append(((w + 3) // 4) * ((h + 3) // 4) * bytes_per_block if bytes_per_block else w * h * bpp)
        if w == 1 and h == 1:
            return out
        w, h = max(1, w // 2), max(1, h // 2)


def test_reader_locates_every_surface_of_reference_written_files(nvtt):
    ref = _ref_or_skip()
    s = nvtt.synth
    img = s.photo_bgra8(100, 60, seed=1, alpha=True)
    cases = [
        # (kwargs for the reference writer, expected block format, input format, bytes per block / per pixel)
        (dict(fmt=ref.Format_RGBA, container=0), -1, 0, None, 4),
        (dict(fmt=ref.Format_BC1, container=0), 1, -1, 8, None),
        (dict(fmt=ref.Format_BC3, container=1), 4, -1, 16, None),
        (dict(fmt=ref.Format_BC5, container=0), 7, -1, 16, None),
        (dict(fmt=ref.Format_BC7, container=1), 11, -1, 16, None),
    ]
    for kw, blockfmt, infmt, bpb, bpp in cases:
        f = ref.process([img], 0, 100, 60, kw["fmt"], 0 if kw["fmt"] == ref.Format_BC7 else 1, header=True, container=kw["container"], mip_filter=0)
        info, surf = nvtt.capi.read_dds(f)
        assert (info.width, info.height, info.depth, info.faceCount, info.textureType) == (100, 60, 1, 1, 0)
        assert info.mipCount == 7 and info.blockFormat == blockfmt and info.inputFormat == infmt
        assert info.headerBytes == (148 if kw["container"] == 1 else 128)
        sizes = _level_sizes(100, 60, bpb, bpp)
        off = info.headerBytes
        for m, n in enumerate(sizes):
            w, h, data = surf[(0, m)]
            assert (w, h) == (max(1, 100 >> m), max(1, 60 >> m)) and data.size == n
            assert np.array_equal(data, f[off:off + n])
            off += n
        assert off == f.size
    # cube map: six faces, face-major
    faces = [s.photo_bgra8(32, 32, seed=10 + i) for i in range(6)]
    f = ref.process(faces, 0, 32, 32, ref.Format_BC1, 1, header=True, texture_type=1, mip_filter=0)
    info, surf = nvtt.capi.read_dds(f)
    assert info.faceCount == 6 and info.textureType == 1 and info.mipCount == 6
    face_bytes = sum(_level_sizes(32, 32, 8))
    assert surf[(3, 0)][2].ctypes.data - surf[(0, 0)][2].ctypes.data == 3 * face_bytes
    assert 128 + 6 * face_bytes == f.size
    # not DDS / truncated
    for bad in (b"\0" * 200, bytes(f[:100]), bytes(f[:128 + 10])):
        with pytest.raises(nvtt.capi.NvttbError):
            nvtt.capi.read_dds(bad)


@pytest.mark.gpu
def test_premade_mip_chain_from_dds_feeds_the_pipeline(nvtt, ctx):
    """nvcompress' DDS input path (tools/compress.cpp:506-557): every (face, mip) surface of an uncompressed .dds goes to
    InputOptions::setMipmapData; the chain is then encoded from the supplied levels.  Here the file is read with OUR reader
    and processed by OUR library, the reference gets the same levels through its own API."""
    import os
    import ctypes as C
    ref = _ref_or_skip()
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libnvtt_b200_harness.so")
    ours = C.CDLL(so)
    ours.ref_process.restype = C.c_long
    ours.ref_process.argtypes = [C.POINTER(ref.RefProcessDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_long]
    img = nvtt.synth.photo_bgra8(64, 48, seed=5, alpha=True)
    # a BGRA8 .dds with a Kaiser mip chain, written by our own writer (Format_RGBA through Compressor::process)
    dds = ref.process([img], 0, 64, 48, ref.Format_RGBA, 1, header=True, mip_filter=2, lib_override=ours)
    assert np.array_equal(dds, ref.process([img], 0, 64, 48, ref.Format_RGBA, 1, header=True, mip_filter=2))
    info, surf = nvtt.capi.read_dds(dds)
    assert info.inputFormat == 0 and info.mipCount == 7
    mips = {(0, m): surf[(0, m)][2].reshape(surf[(0, m)][1], surf[(0, m)][0], 4) for m in range(1, info.mipCount)}
    base = surf[(0, 0)][2].reshape(48, 64, 4)
    for fmt in (ref.Format_BC1, ref.Format_BC3):
        a = ref.process([base], 0, 64, 48, fmt, 1, header=True, mip_filter=0, user_mips=mips, lib_override=ours)
        b = ref.process([base], 0, 64, 48, fmt, 1, header=True, mip_filter=0, user_mips=mips)
        assert a.size == b.size and np.array_equal(a, b), fmt
