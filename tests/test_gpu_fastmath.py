"""The opt-in fast-math build (libnvtt_b200_fastmath.so, NVTT_B200_FASTMATH=1): same C ABI, FMA contraction allowed, OUTSIDE the
bit-exact parity contract.  What is checked: it is a different build (nvttb_build_variant), the strict
library is untouched by its presence, and its quality is the strict library's to within 0.05 dB on BC1 / BC3 / BC5 / BC7."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

_CHILD = r"""
import json, math, sys
sys.path.insert(0, %r)
import numpy as np
import nvtt_b200_loader
m = nvtt_b200_loader.load()
ctx = m.Context(0)
out = {"variant": m.lib().nvttb_build_variant().decode(), "psnr": {}, "sha": {}}
import hashlib
w = h = 256
src = m.synth.photo_bgra8(w, h, seed=99, alpha=True)
for name, fmt, q, bs in (("bc1_production", m.Format_BC1, 2, 8), ("bc1_normal", m.Format_BC1, 1, 8), ("bc3_normal", m.Format_BC3, 1, 16),
                         ("bc5_normal", m.Format_BC5, 1, 16), ("bc7", m.Format_BC7, 1, 16)):
    d = m.make_process_desc(m.InputFormat_BGRA_8UB, w, h, fmt, q, mip_filter=m.MipmapFilter_Kaiser if fmt == m.Format_BC3 else m.MipmapFilter_Box)
    data = ctx.process_bytes([src], d)
    lvl0 = (w // 4) * (h // 4) * bs
    org = m.Surface(ctx); org.set_image(m.InputFormat_BGRA_8UB, w, h, src)
    dec = m.Surface(ctx); dec.set_image_2d(fmt, w, h, data[:lvl0])
    if fmt == m.Format_BC5:  # two channels: compare red / green only
        a = org.get(); b = dec.get()
        rms = float(np.sqrt(np.mean((a[:2] - b[:2]) ** 2)))
    else:
        rms = org.rms_error(dec)
    out["psnr"][name] = 20.0 * math.log10(1.0 / rms)
    out["sha"][name] = hashlib.sha1(data.tobytes()).hexdigest()
print(json.dumps(out))
"""


def _run(fast):
    env = dict(os.environ)
    env.pop("NVTT_B200_FASTMATH", None)
    if fast:
        env["NVTT_B200_FASTMATH"] = "1"
    r = subprocess.run([sys.executable, "-c", _CHILD % ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_fastmath_build_is_separate_and_as_good():
    strict, fast = _run(False), _run(True)
    assert strict["variant"] == "strict" and fast["variant"] == "fastmath"
    for k, p in strict["psnr"].items():
        assert abs(fast["psnr"][k] - p) < 0.05, (k, p, fast["psnr"][k])
    # it really is different arithmetic: at least one of the cluster-fit formats must differ in some block
    assert any(strict["sha"][k] != fast["sha"][k] for k in strict["sha"]), "fast-math build produced identical bytes everywhere"
