"""Mip chain only (no encode) for ncu: python profiles/prof_mips.py <size> <mip_filter 0 box|1 triangle|2 kaiser>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nvtt_b200_loader
m = nvtt_b200_loader.load()
size, filt = int(sys.argv[1]), int(sys.argv[2])
ctx = m.Context(0)
s = m.Surface(ctx)
s.set_image(0, size, size, m.synth.photo_bgra8(size, size, seed=1234, alpha=True))
n = 0
while s.build_next_mipmap(filt):
    n += 1
ctx.synchronize()
print("levels", n)
