"""CUDA-event time of one level encode (no mips), inputs resident: python profiles/time_level.py <FMT> <quality> <size>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nvtt_b200_loader
m = nvtt_b200_loader.load()
fmt_name, quality, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ctx = m.Context(0)
if fmt_name == "RGBA":
    # Format_RGBA (k_pixel_format): planar fp32 -> packed scanlines; quality selects the layout: 0 BGRA8, 1 R5G6B5, 2 RGBA16F, 3 R11G11B10F
    kw = [dict(), dict(masks=(16, 0xF800, 0x7E0, 0x1F, 0)), dict(sizes=(16, 16, 16, 16), pixel_type=4), dict(sizes=(11, 11, 10, 0), pixel_type=4)][quality]
    d = m.make_pixel_format_desc(size, size, **kw)
    n = int(m.lib().nvttb_pixel_format_level_size(d))
    src = torch.rand(4, size, size, device="cuda")
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        ctx.convert_level_device(d, src.data_ptr(), out.data_ptr(), n)
    ctx.synchronize()
    K = 10
    ctx.timer_start()
    for _ in range(K):
        ctx.convert_level_device(d, src.data_ptr(), out.data_ptr(), n)
    ms = ctx.timer_stop() / K
    print("RGBA layout %d %dx%d: %.3f ms, %.0f GB/s (16 B read + %.1f B written per pixel)" % (quality, size, size, ms, (size * size * 16 + n) / ms / 1e6, n / size / size))
    sys.exit(0)
fmt = getattr(m, "Format_" + fmt_name)
if fmt == m.Format_BC6:
    img = torch.from_numpy(m.synth.hdr_rgba16f(size, size, seed=11).view("uint16").astype("int16")).cuda()
    d = m.make_process_desc(m.InputFormat_RGBA_16F, size, size, fmt, quality, mipmaps=False, pixel_type=5)
else:
    img = torch.from_numpy(m.synth.photo_bgra8(size, size, seed=1234, alpha=True)).cuda()
    d = m.make_process_desc(0, size, size, fmt, quality, mipmaps=False)
n = int(m.lib().nvttb_process_output_size(d))
out = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    ctx.process_to_device([img.data_ptr()], d, out.data_ptr(), n)
ctx.synchronize()
K = 5
ctx.timer_start()
for _ in range(K):
    ctx.process_to_device([img.data_ptr()], d, out.data_ptr(), n)
print("%s q%d %dx%d level: %.3f ms" % (fmt_name, quality, size, size, ctx.timer_stop() / K))
