"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference (oracle/_ref, pinned parity build).
Run in the development container (needs /root/reference to have been built by oracle/build_ref.sh):

    python tests/golden/make_golden.py

Inputs are regenerated from seeds (nvtt_b200.synth), only the reference outputs are stored: BCn bytes verbatim, fp32
image-op results as sha256 of their bytes."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refapi as R  # noqa: E402
from golden_cases import level_cases, pipeline_cases, imageop_cases, make_input  # noqa: E402
from golden_cases import level_cases_v2, decode_cases, quantize_cases, pipeline_cases_v2, pixel_format_cases  # noqa: E402


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    out = {}
    for key, (kind, w, h, fmt, q, am, cw, pt) in level_cases().items():
        img = make_input(kind, w, h, planar=True)
        out[key] = R.compress_level(fmt, q, img, alpha_mode=am, color_weights=cw, pixel_type=pt)
    for key, (kind, w, h, fmt, q, kw) in pipeline_cases().items():
        img = make_input(kind, w, h, planar=False)
        out[key] = R.process([img], 0, w, h, fmt, q, **kw)
    for key, (kind, w, h, wrap, filt, params) in imageop_cases().items():
        img = make_input(kind, w, h, planar=False)
        s = R.Surface(wrap=wrap)
        s.set_image(0, w, h, img)
        s.to_linear(2.2)
        hashes = [sha(s.get())]
        while s.build_next_mipmap(filt, params):
            hashes.append(sha(s.get()))
        s.to_gamma(2.2)
        hashes.append(sha(s.get()))
        out[key] = np.stack(hashes)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    if "--v1" in sys.argv:
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes,", len(out), "vectors")
    v1 = out
    # ---- v2: later additions (v1 is left as committed) ----
    out = {}
    lc1, lc2 = level_cases(), level_cases_v2()
    for key, (kind, w, h, fmt, q, am, cw, pt) in lc2.items():
        img = make_input(kind, w, h, planar=True)
        out[key] = R.compress_level(fmt, q, img, alpha_mode=am, color_weights=cw, pixel_type=pt)
    # ZOH::Utils::FORMAT is a global that the BC6H decoder inherits from the last encode: make it "unsigned" (its initial state)
    R.compress_level(10, 1, np.zeros((4, 4, 4), np.float32), pixel_type=5)
    for key, (src, in_v2, dec) in decode_cases().items():
        kind, w, h, fmt = (lc2 if in_v2 else lc1)[src][:4]
        blocks = out[src] if in_v2 else v1[src]
        out[key] = sha(R.decode_ex(fmt, dec, w, h, blocks))
    import nvtt_b200_loader
    synth = nvtt_b200_loader.load().synth
    for key, (w, h, dither) in quantize_cases().items():
        s = R.Surface()
        s.set_image(0, w, h, synth.photo_bgra8(w, h, seed=w + h, alpha=True))
        s.quantize(0, 5, True, dither)
        s.quantize(1, 6, True, dither)
        s.quantize(2, 3, False, dither)
        s.binarize(3, 0.4, dither)
        out[key] = sha(s.get())
    for key, (kind, w, h, fmt, q, kw) in pipeline_cases_v2().items():
        img = make_input(kind, w, h, planar=False)
        out[key] = R.process([img], 0, w, h, fmt, q, **kw)
    # error metrics of two reference-encoded levels against their source image
    for name, src in (("bc1", "level_bc1_photo_48x40_q1"), ("bc3", "level_bc3_photo_48x40_q2")):
        kind, w, h, fmt = lc1[src][:4]
        planar = make_input(kind, w, h, planar=True)
        rgba = np.ascontiguousarray(np.moveaxis(planar, 0, 2))
        for am in (0, 1):
            out["metric_%s_am%d" % (name, am)] = np.array(R.rms_error(fmt, w, h, v1[src], rgba, am), np.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v2.npz")
    if "--v2" in sys.argv or not os.path.exists(path):
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes,", len(out), "vectors")
    # ---- v3: Format_RGBA layouts; one level through Compressor::process (gamma 1 -> no toLinear / toGamma, no mips, no header):
    # the stream is PixelFormatConverter::compress of the BGRA8 image as floats (c / 255) ----
    out = {}
    for key, (w, h, kw) in pixel_format_cases().items():
        img = synth.photo_bgra8(w, h, seed=w * 3 + h, alpha=True)
        out[key] = R.process([img], 0, w, h, 0, 1, mipmaps=False, gamma=(1.0, 1.0), pixel_masks=kw.get("masks"), pixel_sizes=kw.get("sizes"),
                             pixel_type=kw.get("pixel_type", 0), pitch_alignment=kw.get("pitch_alignment", 0))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v3.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "vectors")


if __name__ == "__main__":
    main()
