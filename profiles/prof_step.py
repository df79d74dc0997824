"""One bench step (configs[1]) for ncu: `ncu ... python profiles/prof_step.py [steps]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nvtt_b200_loader  # noqa: E402

m = nvtt_b200_loader.load()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
size = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ctx = m.Context(0)
col = torch.from_numpy(m.synth.photo_bgra8(size, size, seed=1234, alpha=True)).cuda()
nrm = torch.from_numpy(m.synth.normal_bgra8(size, size, seed=7)).cuda()
d3 = m.make_process_desc(0, size, size, m.Format_BC3, 1, mip_filter=2)
d5 = m.make_process_desc(0, size, size, m.Format_BC5, 1, mip_filter=2, normal_map=True)
n = int(m.lib().nvttb_process_output_size(d3))
o3 = torch.empty(n, dtype=torch.uint8, device="cuda")
o5 = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(steps):
    ctx.process_to_device([col.data_ptr()], d3, o3.data_ptr(), n)
    ctx.process_to_device([nrm.data_ptr()], d5, o5.data_ptr(), n)
ctx.synchronize()
print("done", ctx.launches)
