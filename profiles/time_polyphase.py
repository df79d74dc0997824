"""Per-level timing of the mip filter inside the BC5 4096^2 Kaiser chain (CUDA events around every launch):
NVB_NO_TMA=1 selects the old shared-memory kernel.  usage: python profiles/time_polyphase.py [size]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nvtt_b200_loader
m = nvtt_b200_loader.load()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = m.Context(0)
img = m.synth.normal_bgra8(size, size, seed=7)
d_img = torch.from_numpy(img).cuda()
for maxlevel in (2, 3, 4, -1):
    desc = m.make_process_desc(0, size, size, m.Format_BC5, 1, mip_filter=2, normal_map=True, max_level=maxlevel)
    n = int(m.lib().nvttb_process_output_size(desc))
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.process_to_device([d_img.data_ptr()], desc, out.data_ptr(), n)
    ctx.synchronize()
    best = {}
    for _ in range(5):
        ctx.profile_begin()
        ctx.process_to_device([d_img.data_ptr()], desc, out.data_ptr(), n)
        p = ctx.profile_end()
        for k, v in p.items():
            if k not in best or v["total_ms"] < best[k]["total_ms"]:
                best[k] = v
    print("levels", maxlevel, "NO_TMA" if os.environ.get("NVB_NO_TMA") else "TMA", {k: (round(v["total_ms"], 4), v["launches"]) for k, v in sorted(best.items())}, flush=True)
