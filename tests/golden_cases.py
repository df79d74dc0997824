"""Case tables shared by tests/golden/make_golden.py (writer) and the tests that read tests/golden/golden_v1.npz."""
import numpy as np


def make_input(kind, w, h, planar):
    import nvtt_b200_loader
    s = nvtt_b200_loader.load().synth
    if kind == "photo":
        img = s.photo_bgra8(w, h, seed=1234, alpha=True)
    elif kind == "normal":
        img = s.normal_bgra8(w, h, seed=7)
    elif kind == "adv":
        img = s.adversarial_bgra8(w, h, seed=5)
    elif kind == "hdr":
        f = s.hdr_rgba16f(w, h, seed=11).astype(np.float32)
        return np.ascontiguousarray(f.transpose(2, 0, 1)) if planar else s.hdr_rgba16f(w, h, seed=11)
    elif kind == "dark":
        img = (s.photo_bgra8(w, h, seed=3, alpha=True) // 5).astype(np.uint8)
    else:
        raise ValueError(kind)
    return s.planar_from_bgra8(img) if planar else img


def level_cases():
    """key -> (input kind, w, h, nvtt format, quality, alphaMode, colour weights, pixel type)"""
    c = {}
    for kind in ("photo", "adv", "dark"):
        for (w, h) in ((48, 40), (13, 7)):
            for fmt, name, qs in ((1, "bc1", (0, 1, 2, 3)), (4, "bc3", (0, 1, 2, 3)), (6, "bc4", (0, 1, 2)), (7, "bc5", (0, 1, 2)), (3, "bc2", (0, 1)), (5, "bc3n", (0, 1)), (2, "bc1a", (1,))):
                for q in qs:
                    c["level_%s_%s_%dx%d_q%d" % (name, kind, w, h, q)] = (kind, w, h, fmt, q, 0, (1, 1, 1, 1), 0)
    # BC6H: quality is ignored; pixel type 5 = UnsignedFloat, 4 = Float (signed).
    for kind in ("hdr", "photo"):
        for (w, h) in ((32, 24), (13, 7)):
            for pt in (5, 4):
                c["level_bc6_%s_%dx%d_pt%d" % (kind, w, h, pt)] = (kind, w, h, 10, 1, 0, (1, 1, 1, 1), pt)
    for kind in ("photo", "adv"):
        for (w, h) in ((24, 16), (13, 7)):
            c["level_bc7_%s_%dx%d" % (kind, w, h)] = (kind, w, h, 11, 1, 0, (1, 1, 1, 1), 0)
    c["level_bc1_photo_48x40_q2_transp_w"] = ("photo", 48, 40, 1, 2, 1, (0.3, 0.59, 0.11, 1.0), 0)
    c["level_bc3_photo_48x40_q1_transp_w"] = ("photo", 48, 40, 4, 1, 1, (0.3, 0.59, 0.11, 1.0), 0)
    return c


def pipeline_cases():
    """key -> (input kind, w, h, format, quality, kwargs of refapi.process / make_process_desc)"""
    return {
        "pipe_bc1_normal_box_64": ("photo", 64, 64, 1, 1, dict(mip_filter=0, wrap=0)),
        "pipe_bc1_production_box_37x22": ("photo", 37, 22, 1, 2, dict(mip_filter=0)),
        "pipe_bc3_kaiser_mirror_64": ("photo", 64, 64, 4, 1, dict(mip_filter=2, wrap=2)),
        "pipe_bc3_triangle_transp_37x22": ("photo", 37, 22, 4, 1, dict(mip_filter=1, alpha_mode=1)),
        "pipe_bc5_normal_kaiser_64": ("normal", 64, 64, 7, 1, dict(mip_filter=2, normal_map=True)),
        "pipe_bc4_box_repeat_33x16": ("photo", 33, 16, 6, 1, dict(mip_filter=0, wrap=1)),
    }


def imageop_cases():
    """key -> (input kind, w, h, wrap, mipmap filter, params or None): sha256 of every level after toLinear ... toGamma"""
    c = {}
    for (w, h) in ((64, 64), (37, 22), (33, 16), (7, 1)):
        for wrap in (0, 1, 2):
            for filt, params, name in ((0, None, "box"), (1, None, "tri"), (2, (3.0, 4.0, 1.0), "kaiser")):
                c["ops_%s_wrap%d_%dx%d" % (name, wrap, w, h)] = ("photo", w, h, wrap, filt, params)
    return c


# ---- golden_v2.npz: what was added after v1 (more format/quality pairs, decoders, quantisation, containers, metrics) -------
def level_cases_v2():
    c = {}
    for kind in ("photo", "adv", "dark"):
        for (w, h) in ((48, 40), (13, 7)):
            for fmt, name, qs in ((5, "bc3n", (2, 3)), (3, "bc2", (2, 3)), (2, "bc1a", (2, 3)), (7, "bc5", (3,)), (6, "bc4", (3,))):
                for q in qs:
                    c["level_%s_%s_%dx%d_q%d" % (name, kind, w, h, q)] = (kind, w, h, fmt, q, 0, (1, 1, 1, 1), 0)
            for am in (0, 1):
                c["level_bc3rgbm_%s_%dx%d_am%d" % (kind, w, h, am)] = (kind, w, h, 12, 1, am, (1, 1, 1, 1), 0)
    return c


def decode_cases():
    """key -> (level case whose reference blocks are decoded, is it in v2, nvtt::Decoder)"""
    c = {}
    for name, src, v2 in (("bc1", "level_bc1_photo_48x40_q1", False), ("bc2", "level_bc2_adv_13x7_q1", False), ("bc3", "level_bc3_photo_48x40_q2", False),
                          ("bc3n", "level_bc3n_photo_48x40_q1", False), ("bc4", "level_bc4_dark_13x7_q1", False), ("bc5", "level_bc5_photo_48x40_q1", False),
                          ("bc3rgbm", "level_bc3rgbm_photo_48x40_am0", True)):
        for dec in (0, 1, 2):
            c["decode_%s_dec%d" % (name, dec)] = (src, v2, dec)
    c["decode_bc6_unsigned"] = ("level_bc6_hdr_32x24_pt5", False, 0)
    c["decode_bc7"] = ("level_bc7_adv_24x16", False, 0)  # contains no mode-0 block the reference could not decode
    return c


def quantize_cases():
    """key -> (w, h, dither): 5/6/3-bit quantize of r, g, b (the last without exact end points) + binarize of alpha at 0.4"""
    return {"quantize_%dx%d_dither%d" % (w, h, d): (w, h, d) for (w, h) in ((37, 22), (64, 1100), (300, 3)) for d in (0, 1)}


def pipeline_cases_v2():
    return {
        "pipe_ktx_bc3_kaiser_40": ("photo", 40, 40, 4, 1, dict(mip_filter=2, container=2, header=True)),
        "pipe_ktx_bc5_normal_tri_32": ("normal", 32, 32, 7, 1, dict(mip_filter=1, normal_map=True, container=2, header=True)),
        "pipe_dds10_bc7_box_16x12": ("photo", 16, 12, 11, 1, dict(mip_filter=0, container=1, header=True)),
        "pipe_bc1_color_dither_96x72": ("photo", 96, 72, 1, 1, dict(mip_filter=0, quantization=1)),
        "pipe_bc3_binary_alpha_64x48": ("photo", 64, 48, 4, 1, dict(mip_filter=0, quantization=4, alpha_threshold=40)),
    }


def pixel_format_cases():
    """Format_RGBA layouts (PixelFormatConverter, CompressorRGB.cpp:410-575).  key -> (w, h, kwargs of convert_level)."""
    c = {}
    layouts = {
        "bgra8": dict(),
        "rgba8": dict(masks=(32, 0xFF, 0xFF00, 0xFF0000, 0xFF000000)),
        "rgb8_align4": dict(masks=(24, 0xFF0000, 0xFF00, 0xFF, 0), pitch_alignment=4),
        "r5g6b5": dict(masks=(16, 0xF800, 0x7E0, 0x1F, 0)),
        "a1r5g5b5_align8": dict(masks=(16, 0x7C00, 0x3E0, 0x1F, 0x8000), pitch_alignment=8),
        "l8": dict(masks=(8, 0xFF, 0, 0, 0)),
        "a8": dict(masks=(8, 0, 0, 0, 0xFF)),
        "a2r10g10b10": dict(masks=(32, 0x3FF00000, 0xFFC00, 0x3FF, 0xC0000000)),
        "x4r4g4b4_12bit": dict(masks=(12, 0xF00, 0xF0, 0xF, 0)),
        "sizes_5650": dict(sizes=(5, 6, 5, 0)),
        "sizes_10_10_10_2": dict(sizes=(10, 10, 10, 2)),
        "sizes_r16": dict(sizes=(16, 0, 0, 0)),
        "sizes_3320_7bit_stream": dict(sizes=(3, 2, 2, 0)),
        "uint8888": dict(sizes=(8, 8, 8, 8), pixel_type=2),
        "snorm_zero": dict(sizes=(8, 8, 8, 8), pixel_type=1),
        "rgba16f": dict(sizes=(16, 16, 16, 16), pixel_type=4),
        "rgba32f": dict(sizes=(32, 32, 32, 32), pixel_type=4),
        "rg16f": dict(sizes=(16, 16, 0, 0), pixel_type=4),
        "r32f_align4": dict(sizes=(32, 0, 0, 0), pixel_type=4, pitch_alignment=4),
        "r11g11b10f": dict(sizes=(11, 11, 10, 0), pixel_type=4),
    }
    for name, kw in layouts.items():
        for (w, h) in ((37, 22), (64, 12)):
            c["pixfmt_%s_%dx%d" % (name, w, h)] = (w, h, kw)
    return c
