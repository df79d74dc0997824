"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/_build/liboracle.so (plain-C restatement, oracle/oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "liboracle.so")


class OrcProcessDesc(C.Structure):
    _fields_ = [("inputFormat", C.c_int), ("width", C.c_int), ("height", C.c_int), ("wrapMode", C.c_int),
                ("mipmapFilter", C.c_int), ("generateMipmaps", C.c_int), ("maxLevel", C.c_int),
                ("kaiserWidth", C.c_float), ("kaiserAlpha", C.c_float), ("kaiserStretch", C.c_float),
                ("inputGamma", C.c_float), ("outputGamma", C.c_float),
                ("isNormalMap", C.c_int), ("normalizeMipmaps", C.c_int), ("alphaMode", C.c_int),
                ("format", C.c_int), ("quality", C.c_int), ("colorWeights", C.c_float * 4)]


_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.orc_compress_level.restype = C.c_long
        L.orc_compress_level.argtypes = [ci, ci, ci, ci, ci, vp, vp, vp]
        L.orc_process_face.restype = C.c_long
        L.orc_process_face.argtypes = [C.POINTER(OrcProcessDesc), vp, vp]
        L.orc_set_image.argtypes = [ci, ci, ci, vp, vp]
        L.orc_to_linear.argtypes = [vp, ci, ci, cf]
        L.orc_to_gamma.argtypes = [vp, ci, ci, cf]
        L.orc_box_down.argtypes = [vp, ci, ci, vp]
        L.orc_resize.argtypes = [vp, ci, ci, vp, ci, ci, ci, cf, cf, cf, ci]
        L.orc_next_mipmap.argtypes = [vp, ci, ci, vp, ci, cf, cf, cf, ci, ci]
        L.orc_renormalize.argtypes = [vp, ci, ci]
        _lib = L
    return _lib


class OrcPixelFormatDesc(C.Structure):
    _fields_ = [("pixelType", C.c_int), ("bitcount", C.c_uint),
                ("rmask", C.c_uint), ("gmask", C.c_uint), ("bmask", C.c_uint), ("amask", C.c_uint),
                ("rsize", C.c_uint), ("gsize", C.c_uint), ("bsize", C.c_uint), ("asize", C.c_uint),
                ("pitchAlignment", C.c_int), ("width", C.c_int), ("height", C.c_int)]


def convert_level(planar_rgba, masks=None, sizes=None, pixel_type=0, pitch_alignment=1):
    """Format_RGBA: PixelFormatConverter::compress of one level.  masks = (bitcount, r, g, b, a) or sizes = (r, g, b, a);
    neither = CompressionOptions' default BGRA8."""
    a = np.ascontiguousarray(planar_rgba, dtype=np.float32)
    _, h, w = a.shape
    d = OrcPixelFormatDesc()
    d.pixelType, d.pitchAlignment, d.width, d.height = pixel_type, pitch_alignment, w, h
    if sizes is not None:
        d.rsize, d.gsize, d.bsize, d.asize = sizes
    else:
        d.bitcount, d.rmask, d.gmask, d.bmask, d.amask = masks or (32, 0xFF0000, 0xFF00, 0xFF, 0xFF000000)
        if masks is None:
            d.rsize = d.gsize = d.bsize = d.asize = 8
    L = lib()
    L.orc_convert_level.restype = C.c_long
    L.orc_convert_level.argtypes = [C.POINTER(OrcPixelFormatDesc), C.c_void_p, C.c_void_p]
    n = L.orc_convert_level(C.byref(d), a.ctypes.data, None)
    if n <= 0:
        raise RuntimeError("orc_convert_level: unsupported layout")
    out = np.zeros(n, np.uint8)
    assert L.orc_convert_level(C.byref(d), a.ctypes.data, out.ctypes.data) == n
    return out


def block_bytes(fmt):
    return 8 if fmt in (1, 2, 6) else 16


def compress_level(fmt, quality, planar_rgba, alpha_mode=0, color_weights=None):
    a = np.ascontiguousarray(planar_rgba, dtype=np.float32)
    _, h, w = a.shape
    n = ((w + 3) // 4) * ((h + 3) // 4) * block_bytes(fmt)
    out = np.zeros(n, np.uint8)
    cw = (C.c_float * 4)(*color_weights) if color_weights is not None else None
    r = lib().orc_compress_level(fmt, quality, alpha_mode, w, h, a.ctypes.data, cw, out.ctypes.data)
    if r != n:
        raise RuntimeError("orc_compress_level: unsupported format/quality (%d, %d)" % (fmt, quality))
    return out


def process(images, input_format, w, h, fmt, quality, *, wrap=2, mip_filter=0, mipmaps=True, max_level=-1,
            kaiser=(3.0, 4.0, 1.0), gamma=(2.2, 2.2), normal_map=False, normalize_mipmaps=True, alpha_mode=0,
            color_weights=(1, 1, 1, 1)):
    d = OrcProcessDesc()
    d.inputFormat, d.width, d.height, d.wrapMode, d.mipmapFilter = input_format, w, h, wrap, mip_filter
    d.generateMipmaps, d.maxLevel = int(mipmaps), max_level
    d.kaiserWidth, d.kaiserAlpha, d.kaiserStretch = kaiser
    d.inputGamma, d.outputGamma = gamma
    d.isNormalMap, d.normalizeMipmaps, d.alphaMode, d.format, d.quality = int(normal_map), int(normalize_mipmaps), alpha_mode, fmt, quality
    d.colorWeights = (C.c_float * 4)(*color_weights)
    outs = []
    cap = 2 * ((w + 3) // 4) * ((h + 3) // 4) * 16 + 64
    for im in images:
        im = np.ascontiguousarray(im)
        out = np.zeros(cap, np.uint8)
        n = lib().orc_process_face(C.byref(d), im.ctypes.data, out.ctypes.data)
        if n < 0:
            raise RuntimeError("orc_process_face failed")
        outs.append(out[:n])
    return np.concatenate(outs)


def set_image(fmt, w, h, data):
    out = np.empty((4, h, w), np.float32)
    data = np.ascontiguousarray(data)
    lib().orc_set_image(fmt, w, h, data.ctypes.data, out.ctypes.data)
    return out


def to_linear(img, g):
    img = np.ascontiguousarray(img, np.float32).copy()
    lib().orc_to_linear(img.ctypes.data, img.shape[2], img.shape[1], g)
    return img


def to_gamma(img, g):
    img = np.ascontiguousarray(img, np.float32).copy()
    lib().orc_to_gamma(img.ctypes.data, img.shape[2], img.shape[1], g)
    return img


def next_mipmap(img, filt, fwidth, p0, p1, wrap, alpha_mode=0):
    img = np.ascontiguousarray(img, np.float32)
    _, h, w = img.shape
    out = np.empty((4, max(1, h // 2), max(1, w // 2)), np.float32)
    lib().orc_next_mipmap(img.ctypes.data, w, h, out.ctypes.data, filt, fwidth, p0, p1, wrap, alpha_mode)
    return out


def resize(img, dw, dh, kind, fwidth, p0, p1, wrap):
    img = np.ascontiguousarray(img, np.float32)
    _, h, w = img.shape
    out = np.empty((4, dh, dw), np.float32)
    lib().orc_resize(img.ctypes.data, w, h, out.ctypes.data, dw, dh, kind, fwidth, p0, p1, wrap)
    return out


def renormalize(img):
    img = np.ascontiguousarray(img, np.float32).copy()
    lib().orc_renormalize(img.ctypes.data, img.shape[2], img.shape[1])
    return img
