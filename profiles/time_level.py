"""CUDA-event time of one level encode (no mips), inputs resident: python profiles/time_level.py <FMT> <quality> <size>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nvtt_b200_loader
m = nvtt_b200_loader.load()
fmt_name, quality, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
fmt = getattr(m, "Format_" + fmt_name)
ctx = m.Context(0)
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
