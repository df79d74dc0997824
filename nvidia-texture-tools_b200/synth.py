"""Seeded synthetic inputs for tests and bench.  numpy only; BGRA8 unless stated."""
import numpy as np


def _fields(w, h, seed):
    rng = np.random.default_rng(seed)
    y = np.arange(h, dtype=np.float32)[:, None]
    x = np.arange(w, dtype=np.float32)[None, :]
    return rng, x, y


def photo_bgra8(w, h, seed=1234, alpha=False):
    """S1 'photo-like' (alpha=255) or S2 'alpha' (ramp + noise, 10% exact 0/255 runs)."""
    rng, x, y = _fields(w, h, seed)
    out = np.empty((h, w, 4), np.uint8)
    periods = [(37.0, 211.0, 89.0), (53.0, 131.0, 173.0), (71.0, 97.0, 199.0)]
    for c, (p0, p1, p2) in enumerate(periods):
        f = 128.0 + 100.0 * (np.sin(x / p0 + c) * np.cos(y / p1 + 2 * c) * 0.6 + 0.4 * np.sin((x + y) / p2))
        f = f + rng.normal(0.0, 12.0, (h, w)).astype(np.float32)
        out[..., 2 - c] = np.clip(f, 0, 255).astype(np.uint8)  # memory order B,G,R,A
    if alpha:
        a = (x / max(w - 1, 1)) * 255.0 + rng.normal(0.0, 8.0, (h, w)).astype(np.float32)
        a = np.clip(a, 0, 255)
        runs = rng.random((h, (w + 15) // 16)) < 0.10
        runs = np.repeat(runs, 16, axis=1)[:, :w]
        hi = rng.random((h, (w + 15) // 16)) < 0.5
        hi = np.repeat(hi, 16, axis=1)[:, :w]
        a = np.where(runs, np.where(hi, 255.0, 0.0), a)
        out[..., 3] = a.astype(np.uint8)
    else:
        out[..., 3] = 255
    return out


def normal_bgra8(w, h, seed=7):
    """S3: analytic bumps, renormalised, packed to u8 (x->R, y->G, z->B)."""
    rng, x, y = _fields(w, h, seed)
    nx = 0.5 * np.sin(x / 23.0) * np.ones_like(y) + rng.normal(0, 0.02, (h, w)).astype(np.float32)
    ny = 0.5 * np.cos(y / 31.0) * np.ones_like(x) + rng.normal(0, 0.02, (h, w)).astype(np.float32)
    nz = np.sqrt(np.maximum(1.0 - nx * nx - ny * ny, 0.0))
    l = np.sqrt(nx * nx + ny * ny + nz * nz)
    out = np.empty((h, w, 4), np.uint8)
    out[..., 2] = np.clip((nx / l * 0.5 + 0.5) * 255.0 + 0.5, 0, 255).astype(np.uint8)
    out[..., 1] = np.clip((ny / l * 0.5 + 0.5) * 255.0 + 0.5, 0, 255).astype(np.uint8)
    out[..., 0] = np.clip((nz / l * 0.5 + 0.5) * 255.0 + 0.5, 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


def hdr_rgba16f(w, h, seed=11):
    """S4: fp16 exp(N(0,1.5)) * smooth field, 1% > 1000, unsigned, no NaN/Inf."""
    rng, x, y = _fields(w, h, seed)
    base = (1.0 + 0.5 * np.sin(x / 41.0) * np.cos(y / 67.0)).astype(np.float32)
    out = np.empty((h, w, 4), np.float16)
    for c in range(3):
        v = np.exp(rng.normal(0.0, 1.5, (h, w))).astype(np.float32) * base
        hot = rng.random((h, w)) < 0.01
        v = np.where(hot, v * 1000.0 + 1000.0, v)
        out[..., c] = np.minimum(v, 60000.0).astype(np.float16)
    out[..., 3] = np.float16(1.0)
    return out


def adversarial_bgra8(w, h, seed=5):
    """S5: uniform random texels with flat / two-colour 4x4 tiles sprinkled in."""
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    bh, bw = (h + 3) // 4, (w + 3) // 4
    kind = rng.random((bh, bw))
    flat = np.repeat(np.repeat(rng.integers(0, 256, (bh, bw, 4), dtype=np.uint8), 4, 0), 4, 1)[:h, :w]
    flat2 = np.repeat(np.repeat(rng.integers(0, 256, (bh, bw, 4), dtype=np.uint8), 4, 0), 4, 1)[:h, :w]
    pick = rng.random((h, w)) < 0.5
    two = np.where(pick[..., None], flat, flat2)
    k = np.repeat(np.repeat(kind, 4, 0), 4, 1)[:h, :w]
    out = np.where((k < 0.1)[..., None], flat, out)
    out = np.where(((k >= 0.1) & (k < 0.2))[..., None], two, out)
    return np.ascontiguousarray(out)


def planar_from_bgra8(img):
    """BGRA8 [h,w,4] -> planar fp32 RGBA [4,h,w] exactly as Surface::setImage does (x/255.0f)."""
    f = img.astype(np.float32) / np.float32(255.0)
    return np.ascontiguousarray(np.stack([f[..., 2], f[..., 1], f[..., 0], f[..., 3]]))
