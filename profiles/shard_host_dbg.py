import sys, os, threading, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import nvtt_b200_loader
nvtt = nvtt_b200_loader.load()
mode = sys.argv[1] if len(sys.argv) > 1 else "pageable"
w, h, world, chunk = 256, 256, 4, 16
L = nvtt.lib()
ctx = nvtt.Context(0)
img = nvtt.synth.photo_bgra8(w, h, seed=21, alpha=True)
fmt, q = nvtt.Format_BC1, 2
whole = ctx.process_bytes([img], nvtt.make_process_desc(0, w, h, fmt, q))
d0 = nvtt.make_process_desc(0, w, h, fmt, q, band_index=0, band_count=world, band_chunk_rows=chunk)
xbytes = int(L.nvttb_process_exchange_size(C.byref(d0)))
xchg = C.c_void_p()
ctx._ck(L.nvttb_device_alloc(ctx.h, xbytes, C.byref(xchg)))
if mode == "pinned":
    host_t = torch.zeros(whole.size, dtype=torch.uint8).pin_memory(); host = host_t.numpy()
    img_t = torch.from_numpy(img).pin_memory(); img = img_t.numpy()
else:
    host = np.zeros(whole.size, np.uint8)
ctxs = [nvtt.Context(0) for _ in range(world)]
for b in range(world):
    ctxs[b].process_prepare(nvtt.make_process_desc(0, w, h, fmt, q, band_index=b, band_count=world, band_chunk_rows=chunk, band_output_in_place=True, band_exchange=xchg.value, band_sequence=1))
errs = []
times = {}
def run(b):
    try:
        t0 = time.time()
        d = nvtt.make_process_desc(0, w, h, fmt, q, band_index=b, band_count=world, band_chunk_rows=chunk, band_output_in_place=True, band_exchange=xchg.value, band_sequence=1)
        ctxs[b].process_shard([img], d, None, host.ctypes.data)
        times[b] = time.time() - t0
    except Exception as e:
        errs.append((b, e)); times[b] = time.time() - t0
th = [threading.Thread(target=run, args=(b,)) for b in range(world)]
for t in th: t.start()
for t in th: t.join()
print(mode, os.environ.get("CUDA_MODULE_LOADING"), "errs", errs, "times", times, "equal", bool(np.array_equal(host, whole)), flush=True)
