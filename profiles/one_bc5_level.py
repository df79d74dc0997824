"""One 4096^2 BC5 Kaiser chain limited to `levels` levels, for ncu (kernel name filter picks the launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nvtt_b200_loader
m = nvtt_b200_loader.load()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = m.Context(0)
img = m.synth.normal_bgra8(size, size, seed=7)
d_img = torch.from_numpy(img).cuda()
desc = m.make_process_desc(0, size, size, m.Format_BC5, 1, mip_filter=2, normal_map=True, max_level=levels)
n = int(m.lib().nvttb_process_output_size(desc))
out = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    ctx.process_to_device([d_img.data_ptr()], desc, out.data_ptr(), n)
ctx.synchronize()
