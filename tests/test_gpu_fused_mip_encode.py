"""The opt-in fused kernel (NVTT_B200_FUSED_MIP_ENCODE=1): k_polyphase_tma<W, true> filters a 2:1 mip level and block-encodes
the tile it has just built (BC4 / BC5 quick alpha blocks) - north_star's "separable mip filters ... fused with the next
level's block encode".  Its chains must be byte-identical to the reference (and therefore to the unfused path)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

_CHILD = r"""
import json, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import nvtt_b200_loader, refapi
m = nvtt_b200_loader.load()
ctx = m.Context(0)
out = {}
cases = [("bc5_normal_kaiser_1100x1004", "normal", 1100, 1004, m.Format_BC5, dict(mip_filter=2, normal_map=True, wrap=1)),
         ("bc5_normal_triangle_1024x512", "normal", 1024, 512, m.Format_BC5, dict(mip_filter=1, normal_map=True, wrap=0)),
         ("bc5_colour_kaiser_1024x1024", "photo", 1024, 1024, m.Format_BC5, dict(mip_filter=2)),
         ("bc4_colour_kaiser_linear_1100x1004", "photo", 1100, 1004, m.Format_BC4, dict(mip_filter=2, gamma=(1.0, 1.0), wrap=2)),
         ("bc4_fastest_kaiser_1024x1024", "photo", 1024, 1024, m.Format_BC4, dict(mip_filter=2, quality=0))]
for name, kind, w, h, fmt, kw in cases:
    kw = dict(kw)
    q = kw.pop("quality", 1)
    img = m.synth.normal_bgra8(w, h, seed=3) if kind == "normal" else m.synth.photo_bgra8(w, h, seed=5, alpha=True)
    launches0 = ctx.launches
    got = ctx.process_bytes([img], m.make_process_desc(0, w, h, fmt, q, **kw))
    launches = ctx.launches - launches0
    want = refapi.process([img], 0, w, h, fmt, q, **kw)
    out[name] = {"identical": bool(got.size == want.size and np.array_equal(got, want)), "launches": int(launches)}
print(json.dumps(out))
"""


def _run(fused):
    env = dict(os.environ)
    env.pop("NVTT_B200_FUSED_MIP_ENCODE", None)
    if fused:
        env["NVTT_B200_FUSED_MIP_ENCODE"] = "1"
    r = subprocess.run([sys.executable, "-c", _CHILD % (ROOT, os.path.join(ROOT, "oracle"))], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_fused_mip_encode_is_bit_exact():
    plain, fused = _run(False), _run(True)
    for name in plain:
        assert plain[name]["identical"], name
        assert fused[name]["identical"], name
        # the fused path really ran: the big levels' alpha-block launches are gone
        assert fused[name]["launches"] < plain[name]["launches"], (name, fused[name], plain[name])
