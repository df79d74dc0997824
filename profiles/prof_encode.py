"""One level encode (no mips) of a given format/quality/size for ncu:
    ncu ... python profiles/prof_encode.py <BC1|BC3|BC6|BC7|...> <quality> <size> [reps]
Inputs are resident in HBM; the planar fp32 level is built once by the Surface path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nvtt_b200_loader  # noqa: E402

m = nvtt_b200_loader.load()
fmt_name, quality, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
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
for _ in range(reps):
    ctx.process_to_device([img.data_ptr()], d, out.data_ptr(), n)
ctx.synchronize()
print("done", ctx.launches)
