"""Time of one 2:1 mip filter launch (k_polyphase_tma) through nvtt::Surface::buildNextMipmap, per filter (developer tool)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nvtt_b200_loader  # noqa: E402
m = nvtt_b200_loader.load()
ctx = m.Context(0)
for size in (4096, 8192):
    img = m.synth.photo_bgra8(size, size, seed=1, alpha=True)
    for name, filt in (("triangle", 1), ("kaiser", 2)):
        best = 1e9
        for rep in range(5):
            s = m.Surface(ctx, wrap=m.WrapMode_Mirror)
            s.set_image(0, size, size, img)
            ctx.synchronize()
            ctx.profile_begin()
            s.build_next_mipmap(filt)
            prof = ctx.profile_end()
            t = sum(v["total_ms"] for k, v in prof.items() if "polyphase" in k)
            best = min(best, t)
            del s
        gb = size * size * 20 / 1e9  # 16 B read + 4 B written per source texel
        print("%d^2 -> %d^2 %-8s %.4f ms  %.0f GB/s" % (size, size // 2, name, best, gb / (best / 1e3)), flush=True)
