"""Device timing of the BCn decoders (Surface::setImage2D) and rmsError: achieved HBM GB/s against MEASURED_PEAKS.json.
Algorithmic bytes per texel: block bytes / 16 in + 16 out (planar fp32 RGBA) for the decode; 32 in for rmsError."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import nvtt_b200_loader
m = nvtt_b200_loader.load()
ctx = m.Context(0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
size = 4096
rng = np.random.default_rng(1)
planar = m.synth.planar_from_bgra8(m.synth.photo_bgra8(size, size, seed=3, alpha=True))
ref = m.Surface(ctx)
ref.set_image(m.InputFormat_BGRA_8UB, size, size, m.synth.photo_bgra8(size, size, seed=3, alpha=True))
import ctypes as C
for name, fmt, bs in (("BC1", 1, 8), ("BC3", 4, 16), ("BC5", 7, 16), ("BC6", 10, 16), ("BC7", 11, 16)):
    nb = (size // 4) ** 2
    blocks = rng.integers(0, 256, nb * bs, dtype=np.uint8)
    if name == "BC7":
        blocks[0::16] |= 2  # keep away from the all-zero first byte
    d_blocks = torch.from_numpy(blocks).cuda()
    s = m.Surface(ctx)
    for _ in range(3):
        ctx._ck(ctx.L.nvttb_surface_set_image_2d(s.h, fmt, 0, size, size, C.c_void_p(d_blocks.data_ptr()), m.DEVICE, 0))
    ctx.synchronize()
    K = 20
    ctx.timer_start()
    for _ in range(K):
        ctx._ck(ctx.L.nvttb_surface_set_image_2d(s.h, fmt, 0, size, size, C.c_void_p(d_blocks.data_ptr()), m.DEVICE, 0))
    ms = ctx.timer_stop() / K
    gb = size * size * (bs / 16.0 + 16.0) / 1e9
    print("decode %-4s 4096x4096  %7.3f ms  %8.1f Mpix/s  %7.1f GB/s = %4.1f %% of %.0f GB/s" % (name, ms, size * size / 1e6 / (ms / 1e3), gb / (ms / 1e3), 100 * gb / (ms / 1e3) / peak, peak), flush=True)
s = m.Surface(ctx)
s.set_image(m.InputFormat_BGRA_8UB, size, size, m.synth.photo_bgra8(size, size, seed=4, alpha=True))
import time
ref.rms_error(s)
t0 = time.perf_counter()
for _ in range(20):
    ref.rms_error(s)
dt = (time.perf_counter() - t0) / 20
gb = size * size * 32 / 1e9
print("rmsError 4096x4096 (incl. the D2H of the partial sums and the host sync)  %7.3f ms  %7.1f GB/s" % (dt * 1e3, gb / dt))
