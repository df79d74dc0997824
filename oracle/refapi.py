"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/_ref/libnvtt_ref.so (the unmodified reference
built by oracle/build_ref.sh with the pinned parity flags, plus oracle/ref_harness.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# nvtt enums (src/nvtt/nvtt.h:80-277)
Format_RGB, Format_DXT1, Format_DXT1a, Format_DXT3, Format_DXT5, Format_DXT5n, Format_BC4, Format_BC5 = range(8)
Format_RGBA = Format_RGB
Format_DXT1n, Format_CTX1, Format_BC6, Format_BC7, Format_BC3_RGBM = 8, 9, 10, 11, 12
Format_BC1, Format_BC2, Format_BC3, Format_BC3n = Format_DXT1, Format_DXT3, Format_DXT5, Format_DXT5n
Quality_Fastest, Quality_Normal, Quality_Production, Quality_Highest = range(4)
WrapMode_Clamp, WrapMode_Repeat, WrapMode_Mirror = range(3)
InputFormat_BGRA_8UB, InputFormat_RGBA_16F, InputFormat_RGBA_32F, InputFormat_R_32F = range(4)
MipmapFilter_Box, MipmapFilter_Triangle, MipmapFilter_Kaiser = range(3)
ResizeFilter_Box, ResizeFilter_Triangle, ResizeFilter_Kaiser, ResizeFilter_Mitchell = range(4)
AlphaMode_None, AlphaMode_Transparency, AlphaMode_Premultiplied = range(3)
PixelType_UnsignedNorm, PixelType_Float, PixelType_UnsignedFloat = 0, 4, 5
TextureType_2D, TextureType_Cube, TextureType_3D, TextureType_Array = range(4)
Container_DDS, Container_DDS10, Container_KTX = range(3)


class RefProcessDesc(C.Structure):
    _fields_ = [
        ("inputFormat", C.c_int), ("textureType", C.c_int),
        ("width", C.c_int), ("height", C.c_int), ("faces", C.c_int),
        ("wrapMode", C.c_int), ("mipmapFilter", C.c_int), ("generateMipmaps", C.c_int), ("maxLevel", C.c_int),
        ("kaiserWidth", C.c_float), ("kaiserAlpha", C.c_float), ("kaiserStretch", C.c_float),
        ("inputGamma", C.c_float), ("outputGamma", C.c_float),
        ("isNormalMap", C.c_int), ("convertToNormalMap", C.c_int), ("normalizeMipmaps", C.c_int),
        ("alphaMode", C.c_int),
        ("format", C.c_int), ("quality", C.c_int), ("pixelType", C.c_int),
        ("colorWeights", C.c_float * 4),
        ("outputHeader", C.c_int), ("container", C.c_int), ("threads", C.c_int),
        ("quantization", C.c_int), ("alphaThreshold", C.c_int),
        ("pixelFormatMode", C.c_int),
        ("bitcount", C.c_uint), ("rmask", C.c_uint), ("gmask", C.c_uint), ("bmask", C.c_uint), ("amask", C.c_uint),
        ("rsize", C.c_int), ("gsize", C.c_int), ("bsize", C.c_int), ("asize", C.c_int),
        ("pitchAlignment", C.c_int),
        ("maxExtent", C.c_int), ("roundMode", C.c_int), ("userMips", C.c_int), ("mipImages", C.c_void_p),
    ]


_lib = None


def available(fast=False):
    return os.path.exists(os.path.join(_HERE, "_ref", "libnvtt_ref_fast.so" if fast else "libnvtt_ref.so"))


def lib(fast=False):
    """Load the pinned (parity) or fast (timing-only) reference build."""
    global _lib
    if fast:
        return _load("libnvtt_ref_fast.so")
    if _lib is None:
        _lib = _load("libnvtt_ref.so")
    return _lib


def _load(name):
    L = C.CDLL(os.path.join(_HERE, "_ref", name))
    L.ref_compress_level.restype = C.c_long
    L.ref_compress_level.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int, C.c_void_p, C.c_long]
    L.ref_process.restype = C.c_long
    L.ref_process.argtypes = [C.POINTER(RefProcessDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_long]
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
    L.ref_surf_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    for n in ("ref_surf_expand_normals", "ref_surf_pack_normals", "ref_surf_normalize_normal_map"):
        getattr(L, n).argtypes = [C.c_void_p]
    L.ref_surf_to_grey_scale.argtypes = [C.c_void_p] + [C.c_float] * 4
    L.ref_surf_to_normal_map.argtypes = [C.c_void_p] + [C.c_float] * 4
    L.ref_decode.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_decode_ex.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_rms_error.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.ref_surf_quantize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ref_surf_binarize.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
    return L


def block_bytes(fmt):
    return 8 if fmt in (Format_DXT1, Format_DXT1a, Format_BC4) else 16


def level_size(fmt, w, h):
    return ((w + 3) // 4) * ((h + 3) // 4) * block_bytes(fmt)


def compress_level(fmt, quality, planar_rgba, alpha_mode=AlphaMode_None, color_weights=None,
                   pixel_type=PixelType_UnsignedNorm, threads=0, fast=False):
    """planar_rgba: float32 [4,h,w] -> np.uint8 BCn bytes of that one level."""
    a = np.ascontiguousarray(planar_rgba, dtype=np.float32)
    _, h, w = a.shape
    n = level_size(fmt, w, h)
    out = np.empty(n, np.uint8)
    cw = None
    if color_weights is not None:
        cw = (C.c_float * 4)(*color_weights)
    r = lib(fast).ref_compress_level(fmt, quality, alpha_mode, w, h, a.ctypes.data, cw, pixel_type, threads,
                                     out.ctypes.data, n)
    if r != n:
        raise RuntimeError("ref_compress_level failed: %d (expected %d)" % (r, n))
    return out


def process(images, input_format, w, h, fmt, quality, *, wrap=WrapMode_Mirror, mip_filter=MipmapFilter_Box,
            mipmaps=True, max_level=-1, kaiser=(3.0, 4.0, 1.0), gamma=(2.2, 2.2), normal_map=False,
            to_normal_map=False, normalize_mipmaps=True, alpha_mode=AlphaMode_None,
            pixel_type=PixelType_UnsignedNorm, color_weights=(1, 1, 1, 1), header=False, container=Container_DDS,
            texture_type=TextureType_2D, threads=0, fast=False, quantization=0, alpha_threshold=127,
            pixel_masks=None, pixel_sizes=None, pitch_alignment=0, max_extent=0, round_mode=0, user_mips=None, lib_override=None):
    """Whole Compressor::process pipeline on the reference; images = list of per-face level-0 arrays.
    pixel_masks = (bitcount, rmask, gmask, bmask, amask) or pixel_sizes = (r, g, b, a) select the Format_RGBA layout."""
    d = RefProcessDesc()
    set_pixel_format(d, pixel_masks, pixel_sizes, pitch_alignment)
    d.quantization, d.alphaThreshold = quantization, alpha_threshold
    d.inputFormat, d.textureType, d.width, d.height, d.faces = input_format, texture_type, w, h, len(images)
    d.wrapMode, d.mipmapFilter, d.generateMipmaps, d.maxLevel = wrap, mip_filter, int(mipmaps), max_level
    d.kaiserWidth, d.kaiserAlpha, d.kaiserStretch = kaiser
    d.inputGamma, d.outputGamma = gamma
    d.isNormalMap, d.convertToNormalMap, d.normalizeMipmaps = int(normal_map), int(to_normal_map), int(normalize_mipmaps)
    d.alphaMode, d.format, d.quality, d.pixelType = alpha_mode, fmt, quality, pixel_type
    d.colorWeights = (C.c_float * 4)(*color_weights)
    d.outputHeader, d.container, d.threads = int(header), container, threads
    imgs = [np.ascontiguousarray(i) for i in images]
    ptrs = (C.c_void_p * len(imgs))(*[i.ctypes.data for i in imgs])
    d.maxExtent, d.roundMode = max_extent, round_mode
    keep = []
    if user_mips:  # {(face, mip): array} for mip >= 1
        levels = max(m for _, m in user_mips) + 1
        arr = (C.c_void_p * (levels * len(imgs)))()
        for (f, m), a in user_mips.items():
            a = np.ascontiguousarray(a)
            keep.append(a)
            arr[m * len(imgs) + f] = a.ctypes.data
        keep.append(arr)
        d.userMips, d.mipImages = levels, C.cast(arr, C.c_void_p)
    L = lib_override if lib_override is not None else lib(fast)
    n = L.ref_process(C.byref(d), ptrs, None, 0)
    if n < 0:
        raise RuntimeError("ref_process failed")
    out = np.empty(n, np.uint8)
    r = L.ref_process(C.byref(d), ptrs, out.ctypes.data, n)
    if r != n:
        raise RuntimeError("ref_process failed on second pass")
    return out


def set_pixel_format(d, pixel_masks=None, pixel_sizes=None, pitch_alignment=0):
    d.pixelFormatMode = 0
    if pixel_masks is not None:
        d.pixelFormatMode = 1
        d.bitcount, d.rmask, d.gmask, d.bmask, d.amask = pixel_masks
    elif pixel_sizes is not None:
        d.pixelFormatMode = 2
        d.rsize, d.gsize, d.bsize, d.asize = pixel_sizes
    d.pitchAlignment = pitch_alignment


class Surface:
    """Thin handle over nvtt::Surface for image-op parity tests."""

    def __init__(self, wrap=WrapMode_Mirror, alpha_mode=AlphaMode_None, normal_map=False):
        self.L = lib()
        self.h = self.L.ref_surf_create(wrap, alpha_mode, int(normal_map))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_surf_destroy(self.h)
            self.h = None

    def set_image(self, input_format, w, h, data):
        data = np.ascontiguousarray(data)
        assert self.L.ref_surf_set_image(self.h, input_format, w, h, data.ctypes.data)

    def get(self):
        w, h = self.L.ref_surf_width(self.h), self.L.ref_surf_height(self.h)
        out = np.empty((4, h, w), np.float32)
        self.L.ref_surf_get(self.h, out.ctypes.data)
        return out

    def to_linear(self, g):
        self.L.ref_surf_to_linear(self.h, g)

    def to_gamma(self, g):
        self.L.ref_surf_to_gamma(self.h, g)

    def build_next_mipmap(self, filt, params=None):
        if params is None:
            return bool(self.L.ref_surf_build_next_mipmap(self.h, filt, 0, 0, 0, 0))
        return bool(self.L.ref_surf_build_next_mipmap(self.h, filt, 1, params[0], params[1], params[2]))

    def resize(self, w, h, filt, params=None):
        if params is None:
            self.L.ref_surf_resize(self.h, w, h, filt, 0, 0, 0, 0)
        else:
            self.L.ref_surf_resize(self.h, w, h, filt, 1, params[0], params[1], params[2])

    def expand_normals(self):
        self.L.ref_surf_expand_normals(self.h)

    def pack_normals(self):
        self.L.ref_surf_pack_normals(self.h)

    def normalize_normal_map(self):
        self.L.ref_surf_normalize_normal_map(self.h)

    def quantize(self, channel, bits, exact_end_points, dither):
        self.L.ref_surf_quantize(self.h, channel, bits, int(exact_end_points), int(dither))

    def binarize(self, channel, threshold, dither):
        self.L.ref_surf_binarize(self.h, channel, threshold, int(dither))

    def to_grey_scale(self, r, g, b, a):
        self.L.ref_surf_to_grey_scale(self.h, r, g, b, a)

    def to_normal_map(self, sm, md, bg, lg):
        self.L.ref_surf_to_normal_map(self.h, sm, md, bg, lg)


def decode_ex(fmt, decoder, w, h, data):
    """Surface::setImage2D(fmt, decoder, ...) of the reference -> planar fp32 [4,h,w]."""
    out = np.empty((4, h, w), np.float32)
    data = np.ascontiguousarray(data)
    assert lib().ref_decode_ex(fmt, decoder, w, h, data.ctypes.data, out.ctypes.data)
    return out


def rms_error(fmt, w, h, blocks, rgba32f, alpha_mode=0):
    """(nvtt::rmsError, nvtt::rmsAlphaError) between an RGBA32F image and the decoded BCn level."""
    a, b = C.c_float(), C.c_float()
    blocks = np.ascontiguousarray(blocks)
    rgba32f = np.ascontiguousarray(rgba32f, dtype=np.float32)
    assert lib().ref_rms_error(fmt, w, h, blocks.ctypes.data, rgba32f.ctypes.data, alpha_mode, C.byref(a), C.byref(b))
    return a.value, b.value


def decode(fmt, w, h, data):
    out = np.empty((4, h, w), np.float32)
    data = np.ascontiguousarray(data)
    assert lib().ref_decode(fmt, w, h, data.ctypes.data, out.ctypes.data)
    return out
