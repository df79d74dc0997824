"""ctypes binding of include/nvtt_b200.h.  Fails loudly when the CUDA library or a GPU is missing — there is no
CPU fallback anywhere in this package."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NVTT_B200_FASTMATH=1 selects the FMA-contracted build (same C ABI, faster, outside the bit-exact parity contract)
FASTMATH = os.environ.get("NVTT_B200_FASTMATH", "0") not in ("", "0")
LIB_PATH = os.environ.get("NVTT_B200_LIB") or os.path.join(_HERE, "lib", "libnvtt_b200_fastmath.so" if FASTMATH else "libnvtt_b200.so")

# nvtt enums (src/nvtt/nvtt.h:80-277 of the reference)
Format_RGB, Format_DXT1, Format_DXT1a, Format_DXT3, Format_DXT5, Format_DXT5n, Format_BC4, Format_BC5 = range(8)
Format_BC6, Format_BC7, Format_BC3_RGBM = 10, 11, 12
Format_BC1, Format_BC2, Format_BC3, Format_BC3n = Format_DXT1, Format_DXT3, Format_DXT5, Format_DXT5n
Quality_Fastest, Quality_Normal, Quality_Production, Quality_Highest = range(4)
WrapMode_Clamp, WrapMode_Repeat, WrapMode_Mirror = range(3)
InputFormat_BGRA_8UB, InputFormat_RGBA_16F, InputFormat_RGBA_32F, InputFormat_R_32F = range(4)
MipmapFilter_Box, MipmapFilter_Triangle, MipmapFilter_Kaiser = range(3)
ResizeFilter_Box, ResizeFilter_Triangle, ResizeFilter_Kaiser, ResizeFilter_Mitchell = range(4)
AlphaMode_None, AlphaMode_Transparency, AlphaMode_Premultiplied = range(3)
PixelType_UnsignedNorm, PixelType_Float, PixelType_UnsignedFloat = 0, 4, 5
HOST, DEVICE = 0, 1

ERRORS = {0: "OK", 1: "Error_Unknown", 2: "Error_InvalidInput", 3: "Error_UnsupportedFeature", 4: "Error_CudaError",
          5: "Error_FileOpen", 6: "Error_FileWrite", 7: "Error_UnsupportedOutputFormat"}


class NvttbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERRORS.get(code, code), msg))
        self.code = code


class EncodeDesc(C.Structure):
    _fields_ = [("format", C.c_int), ("quality", C.c_int), ("alphaMode", C.c_int), ("pixelType", C.c_int),
                ("colorWeights", C.c_float * 4), ("width", C.c_int), ("height", C.c_int), ("applyToGamma", C.c_int),
                ("rgbmThreshold", C.c_float)]


class PixelFormatDesc(C.Structure):
    """NvttbPixelFormatDesc: CompressionOptions' Format_RGBA layout (setPixelFormat / setPixelType / setPitchAlignment)."""
    _fields_ = [("pixelType", C.c_int), ("bitcount", C.c_uint),
                ("rmask", C.c_uint), ("gmask", C.c_uint), ("bmask", C.c_uint), ("amask", C.c_uint),
                ("rsize", C.c_uint), ("gsize", C.c_uint), ("bsize", C.c_uint), ("asize", C.c_uint),
                ("pitchAlignment", C.c_int), ("width", C.c_int), ("height", C.c_int)]


def make_pixel_format_desc(w, h, masks=None, sizes=None, pixel_type=0, pitch_alignment=1):
    """masks = (bitcount, rmask, gmask, bmask, amask) or sizes = (r, g, b, a); neither = the reference's default BGRA8."""
    d = PixelFormatDesc()
    d.pixelType, d.pitchAlignment, d.width, d.height = pixel_type, pitch_alignment, w, h
    if sizes is not None:
        d.bitcount = 0
        d.rsize, d.gsize, d.bsize, d.asize = sizes
    else:
        d.bitcount, d.rmask, d.gmask, d.bmask, d.amask = masks or (32, 0xFF0000, 0xFF00, 0xFF, 0xFF000000)
        if masks is None:
            d.rsize = d.gsize = d.bsize = d.asize = 8
    return d


class ProcessDesc(C.Structure):
    _fields_ = [("inputFormat", C.c_int), ("width", C.c_int), ("height", C.c_int), ("faceCount", C.c_int),
                ("wrapMode", C.c_int), ("mipmapFilter", C.c_int), ("generateMipmaps", C.c_int), ("maxLevel", C.c_int),
                ("kaiserWidth", C.c_float), ("kaiserAlpha", C.c_float), ("kaiserStretch", C.c_float),
                ("inputGamma", C.c_float), ("outputGamma", C.c_float),
                ("isNormalMap", C.c_int), ("convertToNormalMap", C.c_int), ("normalizeMipmaps", C.c_int),
                ("heightFactors", C.c_float * 4), ("bumpFrequencyScale", C.c_float * 4),
                ("alphaMode", C.c_int), ("encode", EncodeDesc), ("firstFace", C.c_int), ("lastFace", C.c_int),
                ("bandIndex", C.c_int), ("bandCount", C.c_int), ("bandOutputInPlace", C.c_int), ("bandChunkRows", C.c_int),
                ("bandExchange", C.c_void_p), ("bandSequence", C.c_uint)]


class DdsInfo(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("depth", C.c_int), ("mipCount", C.c_int), ("faceCount", C.c_int),
                ("arraySize", C.c_int), ("textureType", C.c_int), ("blockFormat", C.c_int), ("inputFormat", C.c_int),
                ("bitsPerPixel", C.c_uint), ("blockBytes", C.c_uint), ("headerBytes", C.c_uint), ("dxgiFormat", C.c_uint),
                ("fourcc", C.c_uint), ("hasAlpha", C.c_int), ("isNormalMap", C.c_int)]


def read_dds(data):
    """nvttb_dds_describe + nvttb_dds_surface on a .dds file image (bytes / uint8 array): returns (info, {(face, mip): (w, h, uint8 view)})."""
    buf = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
    L = lib()
    info = DdsInfo()
    rc = L.nvttb_dds_describe(buf.ctypes.data, buf.size, C.byref(info))
    if rc != 0:
        raise NvttbError(rc, "not a DDS file the reader supports")
    faces = info.faceCount * (info.arraySize if info.textureType == 3 else 1)
    out = {}
    for f in range(faces):
        for m in range(info.mipCount):
            off, n, w, h, d = C.c_size_t(), C.c_size_t(), C.c_int(), C.c_int(), C.c_int()
            assert L.nvttb_dds_surface(C.byref(info), f, m, C.byref(off), C.byref(n), C.byref(w), C.byref(h), C.byref(d)) == 0
            out[(f, m)] = (w.value, h.value, buf[off.value:off.value + n.value])
    return info, out


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_int), ("total_ms", C.c_double), ("max_ms", C.c_double),
                ("total_units", C.c_double), ("max_units", C.c_double)]


EMIT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t)

EXPORTS = [
    "nvttb_device_count", "nvttb_build_variant", "nvttb_context_create", "nvttb_context_destroy", "nvttb_last_error", "nvttb_launch_count",
    "nvttb_synchronize", "nvttb_stream", "nvttb_timer_start", "nvttb_timer_stop", "nvttb_profile_begin", "nvttb_profile_end", "nvttb_level_size", "nvttb_format_supported", "nvttb_encode_level", "nvttb_pixel_format_level_size", "nvttb_convert_level",
    "nvttb_surface_create", "nvttb_surface_destroy", "nvttb_surface_clone", "nvttb_surface_set_wrap_mode",
    "nvttb_surface_set_alpha_mode", "nvttb_surface_set_normal_map", "nvttb_surface_width", "nvttb_surface_height",
    "nvttb_surface_set_image", "nvttb_surface_to_linear", "nvttb_surface_to_gamma", "nvttb_surface_build_next_mipmap",
    "nvttb_surface_resize", "nvttb_surface_expand_normals", "nvttb_surface_normalize_normal_map",
    "nvttb_surface_pack_normals", "nvttb_surface_to_grey_scale", "nvttb_surface_to_normal_map",
    "nvttb_surface_scale_bias", "nvttb_surface_clamp", "nvttb_surface_range", "nvttb_surface_tone_map", "nvttb_surface_to_rgbm",
    "nvttb_surface_binarize", "nvttb_surface_quantize", "nvttb_surface_set_image_2d", "nvttb_rms_error", "nvttb_rms_alpha_error", "nvttb_angular_error", "nvttb_cielab_error",
    "nvttb_surface_download", "nvttb_surface_device_data", "nvttb_surface_encode", "nvttb_process",
    "nvttb_process_to_device", "nvttb_process_output_size", "nvttb_process_mip_count", "nvttb_process_band_slices",
    "nvttb_dds_describe", "nvttb_dds_surface", "nvttb_process_exchange_size", "nvttb_process_prepare", "nvttb_process_shard", "nvttb_process_multi", "nvttb_bind_thread_to_device", "nvttb_host_register", "nvttb_host_unregister",
    "nvttb_process_whole_output_size", "nvttb_device_alloc", "nvttb_device_free", "nvttb_ipc_export", "nvttb_ipc_open", "nvttb_ipc_close",
]

_lib = None


def lib():
    """dlopen the CUDA library (no GPU needed just to load it and look at its symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, ci, cf, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.nvttb_build_variant.restype = C.c_char_p
    L.nvttb_context_create.argtypes = [ci, C.POINTER(vp)]
    L.nvttb_context_destroy.argtypes = [vp]
    L.nvttb_last_error.argtypes = [vp]
    L.nvttb_last_error.restype = C.c_char_p
    L.nvttb_launch_count.argtypes = [vp]
    L.nvttb_launch_count.restype = C.c_uint64
    L.nvttb_synchronize.argtypes = [vp]
    L.nvttb_stream.argtypes = [vp]
    L.nvttb_stream.restype = vp
    L.nvttb_timer_start.argtypes = [vp]
    L.nvttb_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.nvttb_profile_begin.argtypes = [vp]
    L.nvttb_profile_end.argtypes = [vp, C.POINTER(KernelStat), ci, C.POINTER(ci)]
    L.nvttb_level_size.argtypes = [ci, ci, ci]
    L.nvttb_level_size.restype = sz
    L.nvttb_format_supported.argtypes = [ci, ci]
    L.nvttb_encode_level.argtypes = [vp, C.POINTER(EncodeDesc), vp, ci, vp, ci, sz]
    L.nvttb_pixel_format_level_size.argtypes = [C.POINTER(PixelFormatDesc)]
    L.nvttb_pixel_format_level_size.restype = sz
    L.nvttb_convert_level.argtypes = [vp, C.POINTER(PixelFormatDesc), vp, ci, vp, ci, sz]
    L.nvttb_surface_create.argtypes = [vp, C.POINTER(vp)]
    L.nvttb_surface_destroy.argtypes = [vp]
    L.nvttb_surface_clone.argtypes = [vp, C.POINTER(vp)]
    for n in ("set_wrap_mode", "set_alpha_mode", "set_normal_map"):
        getattr(L, "nvttb_surface_" + n).argtypes = [vp, ci]
        getattr(L, "nvttb_surface_" + n).restype = None
    L.nvttb_surface_width.argtypes = [vp]
    L.nvttb_surface_height.argtypes = [vp]
    L.nvttb_surface_set_image.argtypes = [vp, ci, ci, ci, vp, ci]
    L.nvttb_surface_set_image_2d.argtypes = [vp, ci, ci, ci, ci, vp, ci, ci]
    L.nvttb_surface_scale_bias.argtypes = [vp, ci, ci, cf, cf]
    L.nvttb_surface_clamp.argtypes = [vp, ci, cf, cf]
    L.nvttb_surface_range.argtypes = [vp, ci, ci, cf, C.POINTER(cf), C.POINTER(cf)]
    L.nvttb_surface_tone_map.argtypes = [vp, ci]
    L.nvttb_surface_to_rgbm.argtypes = [vp, cf, cf]
    L.nvttb_surface_binarize.argtypes = [vp, ci, cf, ci]
    L.nvttb_surface_quantize.argtypes = [vp, ci, ci, ci, ci]
    L.nvttb_rms_error.argtypes = [vp, vp, C.POINTER(cf)]
    L.nvttb_rms_alpha_error.argtypes = [vp, vp, C.POINTER(cf)]
    L.nvttb_angular_error.argtypes = [vp, vp, C.POINTER(cf)]
    L.nvttb_cielab_error.argtypes = [vp, vp, C.POINTER(cf)]
    L.nvttb_surface_to_linear.argtypes = [vp, cf]
    L.nvttb_surface_to_gamma.argtypes = [vp, cf]
    L.nvttb_surface_build_next_mipmap.argtypes = [vp, ci, ci, cf, C.POINTER(cf), C.POINTER(ci)]
    L.nvttb_surface_resize.argtypes = [vp, ci, ci, ci, ci, cf, C.POINTER(cf)]
    for n in ("expand_normals", "normalize_normal_map", "pack_normals"):
        getattr(L, "nvttb_surface_" + n).argtypes = [vp]
    L.nvttb_surface_to_grey_scale.argtypes = [vp, cf, cf, cf, cf]
    L.nvttb_surface_to_normal_map.argtypes = [vp, cf, cf, cf, cf]
    L.nvttb_surface_download.argtypes = [vp, vp]
    L.nvttb_surface_device_data.argtypes = [vp]
    L.nvttb_surface_device_data.restype = vp
    L.nvttb_surface_encode.argtypes = [vp, C.POINTER(EncodeDesc), vp, ci, sz]
    L.nvttb_process.argtypes = [vp, C.POINTER(ProcessDesc), C.POINTER(vp), ci, EMIT_FN, vp]
    L.nvttb_process_to_device.argtypes = [vp, C.POINTER(ProcessDesc), C.POINTER(vp), ci, vp, sz, C.POINTER(sz)]
    L.nvttb_process_output_size.argtypes = [C.POINTER(ProcessDesc)]
    L.nvttb_process_output_size.restype = sz
    L.nvttb_process_whole_output_size.argtypes = [C.POINTER(ProcessDesc)]
    L.nvttb_process_whole_output_size.restype = sz
    L.nvttb_device_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.nvttb_device_free.argtypes = [vp, vp]
    L.nvttb_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.nvttb_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.nvttb_ipc_close.argtypes = [vp, vp]
    L.nvttb_process_mip_count.argtypes = [C.POINTER(ProcessDesc)]
    L.nvttb_process_band_slices.argtypes = [C.POINTER(ProcessDesc), C.c_int, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(ci)]
    L.nvttb_process_exchange_size.argtypes = [C.POINTER(ProcessDesc)]
    L.nvttb_process_exchange_size.restype = sz
    L.nvttb_dds_describe.argtypes = [vp, sz, C.POINTER(DdsInfo)]
    L.nvttb_dds_surface.argtypes = [C.POINTER(DdsInfo), ci, ci, C.POINTER(sz), C.POINTER(sz), C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
    L.nvttb_process_prepare.argtypes = [vp, C.POINTER(ProcessDesc), ci, ci]
    L.nvttb_process_shard.argtypes = [vp, C.POINTER(ProcessDesc), C.POINTER(vp), ci, vp, vp]
    L.nvttb_process_multi.argtypes = [C.POINTER(vp), ci, C.POINTER(ProcessDesc), C.POINTER(vp), EMIT_FN, vp]
    L.nvttb_bind_thread_to_device.argtypes = [vp]
    L.nvttb_host_register.argtypes = [vp, vp, sz]
    L.nvttb_host_unregister.argtypes = [vp, vp]
    _lib = L
    return L


def make_encode_desc(fmt, quality, w=0, h=0, alpha_mode=AlphaMode_None, color_weights=(1, 1, 1, 1),
                     pixel_type=PixelType_UnsignedNorm, apply_to_gamma=False, rgbm_threshold=0.15):
    d = EncodeDesc()
    d.rgbmThreshold = rgbm_threshold
    d.format, d.quality, d.alphaMode, d.pixelType = fmt, quality, alpha_mode, pixel_type
    d.colorWeights = (C.c_float * 4)(*color_weights)
    d.width, d.height, d.applyToGamma = w, h, int(apply_to_gamma)
    return d


def make_process_desc(input_format, w, h, fmt, quality, *, faces=1, wrap=WrapMode_Mirror, mip_filter=MipmapFilter_Box,
                      mipmaps=True, max_level=-1, kaiser=(3.0, 4.0, 1.0), gamma=(2.2, 2.2), normal_map=False,
                      to_normal_map=False, normalize_mipmaps=True, alpha_mode=AlphaMode_None,
                      pixel_type=PixelType_UnsignedNorm, color_weights=(1, 1, 1, 1), first_face=0, last_face=0, band_index=0,
                      band_count=0, band_output_in_place=False, band_chunk_rows=0, band_exchange=None, band_sequence=0):
    d = ProcessDesc()
    d.bandOutputInPlace = int(band_output_in_place)
    d.bandChunkRows, d.bandExchange, d.bandSequence = band_chunk_rows, band_exchange, band_sequence
    d.inputFormat, d.width, d.height, d.faceCount = input_format, w, h, faces
    d.wrapMode, d.mipmapFilter, d.generateMipmaps, d.maxLevel = wrap, mip_filter, int(mipmaps), max_level
    d.kaiserWidth, d.kaiserAlpha, d.kaiserStretch = kaiser
    d.inputGamma, d.outputGamma = gamma
    d.isNormalMap, d.convertToNormalMap, d.normalizeMipmaps = int(normal_map), int(to_normal_map), int(normalize_mipmaps)
    d.heightFactors = (C.c_float * 4)(0, 0, 0, 1)
    d.bumpFrequencyScale = (C.c_float * 4)(1.0 / 1.875, 0.5 / 1.875, 0.25 / 1.875, 0.125 / 1.875)
    d.alphaMode = alpha_mode
    d.encode = make_encode_desc(fmt, quality, alpha_mode=alpha_mode, color_weights=color_weights, pixel_type=pixel_type)
    d.firstFace, d.lastFace = first_face, last_face
    d.bandIndex, d.bandCount = band_index, band_count
    return d


class Context:
    """One GPU context (stream + tables + scratch).  Raises NvttbError(Error_CudaError) when there is no GPU."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.nvttb_context_create(device, C.byref(h))
        if rc != 0:
            raise NvttbError(rc, "nvttb_context_create(device=%d) failed (no CUDA device?) — no CPU fallback exists" % device)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.nvttb_context_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise NvttbError(rc, self.L.nvttb_last_error(self.h).decode())

    @property
    def launches(self):
        return int(self.L.nvttb_launch_count(self.h))

    @property
    def stream(self):
        return self.L.nvttb_stream(self.h)

    def timer_start(self):
        self._ck(self.L.nvttb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._ck(self.L.nvttb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile_begin(self):
        self._ck(self.L.nvttb_profile_begin(self.h))

    def profile_end(self):
        st = (KernelStat * 16)()
        n = C.c_int(0)
        self._ck(self.L.nvttb_profile_end(self.h, st, 16, C.byref(n)))
        return {st[i].name.decode(): dict(launches=st[i].launches, total_ms=st[i].total_ms, max_ms=st[i].max_ms,
                                          total_units=st[i].total_units, max_units=st[i].max_units) for i in range(n.value)}

    def synchronize(self):
        self._ck(self.L.nvttb_synchronize(self.h))

    def bind_thread(self):
        """Keep the calling thread on the CPUs next to this context's GPU (no-op where sysfs does not expose the topology)."""
        self._ck(self.L.nvttb_bind_thread_to_device(self.h))

    def encode_level(self, fmt, quality, planar_rgba, **kw):
        """planar_rgba float32 [4,h,w] on the host -> np.uint8 BCn bytes."""
        a = np.ascontiguousarray(planar_rgba, dtype=np.float32)
        _, h, w = a.shape
        d = make_encode_desc(fmt, quality, w, h, **kw)
        n = self.L.nvttb_level_size(fmt, w, h)
        out = np.empty(n, np.uint8)
        self._ck(self.L.nvttb_encode_level(self.h, C.byref(d), a.ctypes.data, HOST, out.ctypes.data, HOST, n))
        return out

    def convert_level(self, planar_rgba, **kw):
        """Format_RGBA: planar_rgba float32 [4,h,w] on the host -> np.uint8 scanlines (PixelFormatConverter::compress)."""
        a = np.ascontiguousarray(planar_rgba, dtype=np.float32)
        _, h, w = a.shape
        d = make_pixel_format_desc(w, h, **kw)
        n = self.L.nvttb_pixel_format_level_size(C.byref(d))
        if n == 0:
            raise RuntimeError("unsupported pixel format")
        out = np.empty(n, np.uint8)
        self._ck(self.L.nvttb_convert_level(self.h, C.byref(d), a.ctypes.data, HOST, out.ctypes.data, HOST, n))
        return out

    def convert_level_device(self, desc, d_rgba_ptr, d_out_ptr, cap):
        self._ck(self.L.nvttb_convert_level(self.h, C.byref(desc), d_rgba_ptr, DEVICE, d_out_ptr, DEVICE, cap))

    def encode_level_device(self, desc, d_rgba_ptr, d_out_ptr, cap):
        self._ck(self.L.nvttb_encode_level(self.h, C.byref(desc), d_rgba_ptr, DEVICE, d_out_ptr, DEVICE, cap))

    def process(self, images, desc, location=HOST):
        """Runs the whole pipeline; images = list of numpy arrays (host) or raw device pointers.
        Returns list of (face, mip, w, h, bytes)."""
        out = []

        def _emit(user, face, mip, w, h, d, data, size):
            out.append((face, mip, w, h, np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_uint8)), (size,)).copy()))
            return 1

        cb = EMIT_FN(_emit)
        ptrs, keep = self._image_ptrs(images, location)
        self._ck(self.L.nvttb_process(self.h, C.byref(desc), ptrs, location, cb, None))
        return out

    def process_bytes(self, images, desc, location=HOST):
        parts = [b for (_, _, _, _, b) in self.process(images, desc, location)]
        return np.concatenate(parts) if parts else np.zeros(0, np.uint8)  # a band may own nothing of a small image

    def process_to_device(self, images, desc, d_out_ptr, cap, location=DEVICE):
        ptrs, keep = self._image_ptrs(images, location)
        written = C.c_size_t(0)
        self._ck(self.L.nvttb_process_to_device(self.h, C.byref(desc), ptrs, location, d_out_ptr, cap, C.byref(written)))
        return written.value

    def process_prepare(self, desc, location=HOST, own_output=True):
        """nvttb_process_prepare: size the buffers of a band-local shard call up front (several bands on one GPU)."""
        self._ck(self.L.nvttb_process_prepare(self.h, C.byref(desc), location, int(own_output)))

    def process_shard(self, images, desc, d_out_ptr=None, h_out_ptr=None, location=HOST):
        """nvttb_process_shard: this band's share of one block-row sharded image, whole-chain layout on device and / or host."""
        ptrs, keep = self._image_ptrs(images, location)
        self._ck(self.L.nvttb_process_shard(self.h, C.byref(desc), ptrs, location, d_out_ptr, h_out_ptr))

    @staticmethod
    def _image_ptrs(images, location):
        keep = []
        vals = []
        for im in images:
            if isinstance(im, np.ndarray):
                im = np.ascontiguousarray(im)
                keep.append(im)
                vals.append(im.ctypes.data)
            else:
                vals.append(int(im))
        return (C.c_void_p * len(vals))(*vals), keep


def process_multi(contexts, images, desc):
    """nvttb_process_multi: host images -> whole chain, on all the given contexts (GPUs) of this process.
    Returns list of (face, mip, w, h, bytes)."""
    out = []

    def _emit(user, face, mip, w, h, d, data, size):
        out.append((face, mip, w, h, np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_uint8)), (size,)).copy()))
        return 1

    cb = EMIT_FN(_emit)
    ptrs, keep = Context._image_ptrs(images, HOST)
    hs = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    rc = lib().nvttb_process_multi(hs, len(contexts), C.byref(desc), ptrs, cb, None)
    if rc != 0:
        raise NvttbError(rc, lib().nvttb_last_error(contexts[0].h).decode())
    return out


class Surface:
    """Device-resident nvtt::Surface mirror (subset on the hot path)."""

    def __init__(self, ctx, wrap=WrapMode_Mirror, alpha_mode=AlphaMode_None, normal_map=False):
        self.ctx, self.L = ctx, ctx.L
        h = C.c_void_p()
        ctx._ck(self.L.nvttb_surface_create(ctx.h, C.byref(h)))
        self.h = h
        self.L.nvttb_surface_set_wrap_mode(h, wrap)
        self.L.nvttb_surface_set_alpha_mode(h, alpha_mode)
        self.L.nvttb_surface_set_normal_map(h, int(normal_map))

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.L.nvttb_surface_destroy(self.h)
        self.h = None

    width = property(lambda s: s.L.nvttb_surface_width(s.h))
    height = property(lambda s: s.L.nvttb_surface_height(s.h))

    def set_image(self, input_format, w, h, data):
        data = np.ascontiguousarray(data)
        self._keep = data
        self.ctx._ck(self.L.nvttb_surface_set_image(self.h, input_format, w, h, data.ctypes.data, HOST))

    def set_image_2d(self, fmt, w, h, blocks, decoder=0, bc6_signed=False):
        """Surface::setImage2D: decode a BCn level (host bytes) into this surface."""
        blocks = np.ascontiguousarray(blocks)
        self._keep = blocks
        self.ctx._ck(self.L.nvttb_surface_set_image_2d(self.h, fmt, decoder, w, h, C.c_void_p(blocks.ctypes.data), HOST, int(bc6_signed)))

    def rms_error(self, img):
        """nvtt::rmsError(self as reference, img)."""
        v = C.c_float()
        self.ctx._ck(self.L.nvttb_rms_error(self.h, img.h, C.byref(v)))
        return v.value

    def cielab_error(self, img):
        v = C.c_float()
        self.ctx._ck(self.L.nvttb_cielab_error(self.h, img.h, C.byref(v)))
        return v.value

    def angular_error(self, img):
        v = C.c_float()
        self.ctx._ck(self.L.nvttb_angular_error(self.h, img.h, C.byref(v)))
        return v.value

    def rms_alpha_error(self, img):
        v = C.c_float()
        self.ctx._ck(self.L.nvttb_rms_alpha_error(self.h, img.h, C.byref(v)))
        return v.value

    def get(self):
        out = np.empty((4, self.height, self.width), np.float32)
        self.ctx._ck(self.L.nvttb_surface_download(self.h, out.ctypes.data))
        return out

    def to_linear(self, g):
        self.ctx._ck(self.L.nvttb_surface_to_linear(self.h, g))

    def to_gamma(self, g):
        self.ctx._ck(self.L.nvttb_surface_to_gamma(self.h, g))

    def build_next_mipmap(self, filt, params=None):
        built = C.c_int(0)
        if params is None:
            self.ctx._ck(self.L.nvttb_surface_build_next_mipmap(self.h, filt, 0, 0.0, None, C.byref(built)))
        else:
            p = (C.c_float * 2)(params[1], params[2])
            self.ctx._ck(self.L.nvttb_surface_build_next_mipmap(self.h, filt, 1, params[0], p, C.byref(built)))
        return bool(built.value)

    def resize(self, w, h, filt, params=None):
        if params is None:
            self.ctx._ck(self.L.nvttb_surface_resize(self.h, w, h, filt, 0, 0.0, None))
        else:
            p = (C.c_float * 2)(params[1], params[2])
            self.ctx._ck(self.L.nvttb_surface_resize(self.h, w, h, filt, 1, params[0], p))

    def expand_normals(self):
        self.ctx._ck(self.L.nvttb_surface_expand_normals(self.h))

    def pack_normals(self):
        self.ctx._ck(self.L.nvttb_surface_pack_normals(self.h))

    def normalize_normal_map(self):
        self.ctx._ck(self.L.nvttb_surface_normalize_normal_map(self.h))

    def to_grey_scale(self, r, g, b, a):
        self.ctx._ck(self.L.nvttb_surface_to_grey_scale(self.h, r, g, b, a))

    def to_normal_map(self, sm, md, bg, lg):
        self.ctx._ck(self.L.nvttb_surface_to_normal_map(self.h, sm, md, bg, lg))

    def encode(self, fmt, quality, **kw):
        d = make_encode_desc(fmt, quality, **kw)
        n = self.L.nvttb_level_size(fmt, self.width, self.height)
        out = np.empty(n, np.uint8)
        self.ctx._ck(self.L.nvttb_surface_encode(self.h, C.byref(d), out.ctypes.data, HOST, n))
        return out
