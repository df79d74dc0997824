"""nvttb_process_multi on the GPUs of ONE process (no torchrun): one host thread per GPU, block-row sharded 8192^2 BC1 Production
(band-local front end, exchange over NVLink peer memory, each GPU copies its slices into the one pinned host buffer) and a
cube map dealt out face by face.  Checks byte identity with one GPU and prints the end-to-end times.
usage: python profiles/multi_gpu_inprocess.py [size] [steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import nvtt_b200_loader  # noqa: E402
import bench  # noqa: E402

m = nvtt_b200_loader.load()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = m.lib().nvttb_device_count()
ctxs = [m.Context(i) for i in range(n)]
img = bench.c2_image(m, size)
desc = m.make_process_desc(0, size, size, m.Format_BC1, m.Quality_Production, mip_filter=0)
want = np.concatenate([b[4] for b in ctxs[0].process([img], desc)])
out = {"gpus": n, "size": size}
for k in sorted({1, 2, 4, n} & set(range(1, n + 1))):
    got = np.concatenate([b[4] for b in m.capi.process_multi(ctxs[:k], [img], desc)])
    ok = bool(np.array_equal(got, want))
    t0 = time.perf_counter()
    for _ in range(steps):
        m.capi.process_multi(ctxs[:k], [img], desc)
    dt = (time.perf_counter() - t0) / steps
    out["gpus_%d" % k] = {"identical": ok, "ms_per_image_e2e_pageable_host": dt * 1e3, "mpix_per_s": size * size / 1e6 / dt}
faces = [m.synth.hdr_rgba16f(512, 512, seed=30 + i) for i in range(6)]
dc = m.make_process_desc(m.InputFormat_RGBA_16F, 512, 512, m.Format_BC6, 1, faces=6, pixel_type=m.PixelType_UnsignedFloat)
w6 = ctxs[0].process(faces, dc)
g6 = m.capi.process_multi(ctxs, faces, dc)
out["cube_identical"] = bool(all(np.array_equal(a[4], b[4]) for a, b in zip(w6, g6)))
print(json.dumps(out))
