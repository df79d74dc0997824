"""Seeded random sweep of Compressor::process configurations against the reference: sizes from 1 x 1 to a few hundred texels
(banded and unbanded host input, with and without the side / second encode streams), every cheap format, filters, wrap modes,
gammas, alpha modes, normal maps, cube faces, truncated chains.  Byte-identical chains are required.  (The fixed cases of
test_gpu_parity.py stay; this is the net for the host-side pipeline logic.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    fmts = ["BC1", "BC3", "BC4", "BC5", "BC1a", "BC2", "BC3n"]
    for i in range(n):
        w = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 13, 16, 31, 32, 37, 64, 100, 128, 130, 255, 256, 300, 512]))
        h = int(rng.choice([1, 2, 3, 4, 6, 9, 16, 22, 33, 60, 64, 96, 128, 200, 256, 384]))
        fmt = str(rng.choice(fmts))
        q = int(rng.choice([0, 1, 1, 2])) if fmt in ("BC1", "BC3", "BC2", "BC3n", "BC1a") else int(rng.choice([0, 1]))
        kw = dict(mip_filter=int(rng.integers(0, 3)), wrap=int(rng.integers(0, 3)))
        if rng.random() < 0.25:
            kw["gamma"] = (1.0, 1.0)
        if rng.random() < 0.15:
            kw["gamma"] = (2.2, 1.0)
        if rng.random() < 0.2:
            kw["alpha_mode"] = 1
        if rng.random() < 0.15:
            kw["mipmaps"] = False
        elif rng.random() < 0.2:
            kw["max_level"] = int(rng.integers(1, 5))
        kind = "alpha"
        if fmt in ("BC5", "BC3n") and rng.random() < 0.6:
            kw["normal_map"] = True
            kind = "normal"
            kw.pop("gamma", None)
        faces = 6 if (w == h and rng.random() < 0.15) else 1
        yield i, w, h, fmt, q, kw, kind, faces
    # larger images: nine upload bands, the TMA filter kernel (levels of >= 128 tiles), partial tiles and blocks
    big = [(1100, 1004, "BC3", 1, dict(mip_filter=2, wrap=2), "alpha"), (1024, 640, "BC1", 2, dict(mip_filter=0), "alpha"),
           (1100, 1004, "BC5", 1, dict(mip_filter=2, wrap=0, normal_map=True), "normal"), (2048, 512, "BC4", 1, dict(mip_filter=1, wrap=1), "alpha"),
           (1030, 1026, "BC1", 1, dict(mip_filter=2, wrap=1, alpha_mode=1), "alpha"), (1536, 1024, "BC2", 1, dict(mip_filter=0, gamma=(1.0, 1.0)), "alpha")]
    for k, (w, h, fmt, q, kw, kind) in enumerate(big):
        yield n + k, w, h, fmt, q, kw, kind, 1


def test_random_pipelines_match_reference(nvtt, ref, ctx):
    s = nvtt.synth
    bad = []
    for i, w, h, fmt, q, kw, kind, faces in _cases(90, 20261018):
        imgs = [(s.normal_bgra8(w, h, seed=40 + i + f) if kind == "normal" else s.photo_bgra8(w, h, seed=40 + i + f, alpha=True)) for f in range(faces)]
        name = {"BC1a": "DXT1a"}.get(fmt, fmt)
        f_ours, f_ref = getattr(nvtt, "Format_" + name), getattr(ref, "Format_" + name)
        d = nvtt.make_process_desc(0, w, h, f_ours, q, faces=faces, **kw)
        got = ctx.process_bytes(imgs, d)
        want = ref.process(imgs, 0, w, h, f_ref, q, texture_type=(ref.TextureType_Cube if faces == 6 else ref.TextureType_2D), **kw)
        if fmt == "BC1a" and q == 0:
            continue  # QuickCompress::compressDXT1a reads uninitialised memory for blocks with alpha 0
        if got.size != want.size or not np.array_equal(got, want):
            bad.append((i, w, h, fmt, q, kw, faces))
        # the same chain through the device-resident entry point (other stream choreography: one level-0 launch)
        if (i % 3 == 0 or i >= 90) and faces == 1:
            import torch
            t = torch.from_numpy(np.ascontiguousarray(imgs[0])).cuda()
            n = int(nvtt.lib().nvttb_process_output_size(d))
            out = torch.empty(n, dtype=torch.uint8, device="cuda")
            ctx.process_to_device([t.data_ptr()], d, out.data_ptr(), n)
            ctx.synchronize()
            if not np.array_equal(out.cpu().numpy(), want):
                bad.append(("device", i, w, h, fmt, q, kw))
    assert not bad, "%d configurations differ from the reference: %s" % (len(bad), bad[:5])
