import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
import nvtt_b200_loader
m = nvtt_b200_loader.load()
ctx = m.Context(0)
s = m.synth
img = s.planar_from_bgra8(s.photo_bgra8(52, 36, seed=3, alpha=True))
for fmt in (1, 2, 3, 4, 5, 6, 7, 11, 12):
    for q in (1, 2, 3):
        if m.lib().nvttb_format_supported(fmt, q):
            ctx.encode_level(fmt, q, img)
hdr = np.ascontiguousarray(np.moveaxis(s.hdr_rgba16f(52, 36, seed=5).astype(np.float32), 2, 0))
b6 = ctx.encode_level(10, 1, hdr, pixel_type=5)
b7 = ctx.encode_level(11, 1, img)
su = m.Surface(ctx)
su.set_image_2d(10, 52, 36, b6)
su.set_image_2d(11, 52, 36, b7)
su.set_image(0, 37, 22, s.photo_bgra8(37, 22, seed=1, alpha=True))
ctx._ck(ctx.L.nvttb_surface_quantize(su.h, 0, 5, 1, 1))
ctx._ck(ctx.L.nvttb_surface_binarize(su.h, 3, 0.5, 1))
d = m.make_process_desc(0, 64, 48, 11, 1, mip_filter=2)
ctx.process_bytes([s.photo_bgra8(64, 48, seed=9, alpha=True)], d)
# Format_RGBA layouts: x4 kernel, per-pixel kernel (odd width, 24-bit, pitch padding) and the per-scanline bit stream
img4 = s.planar_from_bgra8(s.photo_bgra8(64, 20, seed=4, alpha=True))
for kw in (dict(), dict(masks=(16, 0xF800, 0x7E0, 0x1F, 0)), dict(masks=(8, 0xFF, 0, 0, 0)), dict(sizes=(16, 16, 16, 16), pixel_type=4),
           dict(sizes=(32, 32, 32, 32), pixel_type=4), dict(masks=(24, 0xFF0000, 0xFF00, 0xFF, 0), pitch_alignment=4),
           dict(sizes=(11, 11, 10, 0), pixel_type=4), dict(masks=(12, 0xF00, 0xF0, 0xF, 0))):
    ctx.convert_level(img4, **kw)
    ctx.convert_level(img, **kw)
ref_s = m.Surface(ctx)
ref_s.set_image(0, 52, 36, s.photo_bgra8(52, 36, seed=3, alpha=True))
dec = m.Surface(ctx)
dec.set_image_2d(1, 52, 36, ctx.encode_level(1, 1, img))
ref_s.cielab_error(dec)
print("done")
