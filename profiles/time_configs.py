"""Device-resident timing of individual BASELINE configs (developer tool; bench.py is the contract)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nvtt_b200_loader
m = nvtt_b200_loader.load()
ctx = m.Context(0)
cases = {
    "c0_bc1_normal_2048_box": (2048, m.Format_BC1, 1, dict(mip_filter=0)),
    "c2_bc1_production_8192_box": (8192, m.Format_BC1, 2, dict(mip_filter=0)),
    "bc1_normal_8192_box": (8192, m.Format_BC1, 1, dict(mip_filter=0)),
    "bc1_fastest_8192_box": (8192, m.Format_BC1, 0, dict(mip_filter=0)),
    "c1_bc3_normal_4096_kaiser": (4096, m.Format_BC3, 1, dict(mip_filter=2)),
    "c1_bc5_normal_4096_kaiser": (4096, m.Format_BC5, 1, dict(mip_filter=2, normal_map=True)),
    "bc5_production_2048_box": (2048, m.Format_BC5, 2, dict(mip_filter=0, normal_map=True)),
    "c4_bc6h_face_2048_fp16_box": (2048, m.Format_BC6, 1, dict(mip_filter=0, pixel_type=5)),
    "c3_bc7_1024_box": (1024, m.Format_BC7, 1, dict(mip_filter=0)),
    "c3_bc7_2048_box": (2048, m.Format_BC7, 1, dict(mip_filter=0)),
    "c3_bc7_4096_box": (4096, m.Format_BC7, 1, dict(mip_filter=0)),
    "bc7_8192_box": (8192, m.Format_BC7, 1, dict(mip_filter=0)),
    "c4_bc6h_cube_6x2048_fp16_box": (2048, m.Format_BC6, 1, dict(mip_filter=0, pixel_type=5, faces=6)),
}
sel = sys.argv[1:] or list(cases)
for name in sel:
    size, fmt, q, kw = cases[name]
    kw = dict(kw)
    faces = kw.pop("faces", 1)
    if fmt == m.Format_BC6:
        imgs = [torch.from_numpy(m.synth.hdr_rgba16f(size, size, seed=11 + f).view("uint16").astype("int16")).cuda() for f in range(faces)]
        img = imgs[0]
        d = m.make_process_desc(m.InputFormat_RGBA_16F, size, size, fmt, q, faces=faces, **kw) if faces > 1 else m.make_process_desc(m.InputFormat_RGBA_16F, size, size, fmt, q, **kw)
    else:
        img = torch.from_numpy(m.synth.photo_bgra8(size, size, seed=1234, alpha=True)).cuda()
        d = m.make_process_desc(0, size, size, fmt, q, **kw)
    if not m.lib().nvttb_format_supported(fmt, q):
        print(name, "not supported yet")
        continue
    ptrs = [i.data_ptr() for i in imgs] if (fmt == m.Format_BC6 and faces > 1) else [img.data_ptr()]
    n = int(m.lib().nvttb_process_output_size(d))
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        ctx.process_to_device(ptrs, d, out.data_ptr(), n)
    ctx.synchronize()
    ctx.timer_start()
    K = 5 if size * size * (50 if fmt == m.Format_BC7 else 1) < (1 << 27) else 1
    for _ in range(K):
        ctx.process_to_device(ptrs, d, out.data_ptr(), n)
    ms = ctx.timer_stop() / K
    ctx.profile_begin()
    ctx.process_to_device(ptrs, d, out.data_ptr(), n)
    prof = ctx.profile_end()
    print("%-32s %8.3f ms  %9.1f Mpix/s   %s" % (name, ms, faces * size * size / 1e6 / (ms / 1e3),
          {k: round(v["total_ms"], 3) for k, v in prof.items()}), flush=True)
