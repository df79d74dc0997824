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
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "vectors")


if __name__ == "__main__":
    main()
