"""GPU parity against the committed golden fixtures (reference outputs) and against the plain-C oracle restatement."""
import hashlib
import os

import numpy as np
import pytest

import golden_cases as G

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def test_gpu_levels_match_golden(nvtt, ctx, golden):
    for key, (kind, w, h, fmt, q, am, cw, pt) in G.level_cases().items():
        img = G.make_input(kind, w, h, planar=True)
        got = ctx.encode_level(fmt, q, img, alpha_mode=am, color_weights=cw, pixel_type=pt)
        assert np.array_equal(got, golden[key]), key


def test_gpu_pipeline_matches_golden(nvtt, ctx, golden):
    for key, (kind, w, h, fmt, q, kw) in G.pipeline_cases().items():
        img = G.make_input(kind, w, h, planar=False)
        d = nvtt.make_process_desc(0, w, h, fmt, q, **kw)
        assert np.array_equal(ctx.process_bytes([img], d), golden[key]), key


def test_gpu_image_ops_match_golden(nvtt, ctx, golden):
    for key, (kind, w, h, wrap, filt, params) in G.imageop_cases().items():
        img = G.make_input(kind, w, h, planar=False)
        s = nvtt.Surface(ctx, wrap=wrap)
        s.set_image(0, w, h, img)
        s.to_linear(2.2)
        hashes = [_sha(s.get())]
        while s.build_next_mipmap(filt, params):
            hashes.append(_sha(s.get()))
        s.to_gamma(2.2)
        hashes.append(_sha(s.get()))
        assert np.array_equal(np.stack(hashes), golden[key]), key


def test_gpu_matches_oracle_restatement(nvtt, ctx):
    import oracleapi
    if not oracleapi.available():
        pytest.fail("oracle/_build/liboracle.so missing: run __graft_entry__.build()")
    s = nvtt.synth
    for (w, h) in ((128, 128), (61, 35)):
        col = s.photo_bgra8(w, h, seed=21, alpha=True)
        nrm = s.normal_bgra8(w, h, seed=22)
        for fmt, q, img, kw in ((1, 1, col, dict(mip_filter=0)), (1, 2, col, dict(mip_filter=2, alpha_mode=1)),
                                (4, 1, col, dict(mip_filter=2)), (7, 1, nrm, dict(mip_filter=2, normal_map=True)),
                                (6, 1, col, dict(mip_filter=1, wrap=0))):
            d = nvtt.make_process_desc(0, w, h, fmt, q, **kw)
            got = ctx.process_bytes([img], d)
            want = oracleapi.process([img], 0, w, h, fmt, q, **kw)
            assert np.array_equal(got, want), (w, h, fmt, q, kw)
