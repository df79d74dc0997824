"""BASELINE configs[2]: ONE 8192x8192 RGBA8 image -> BC1 Quality_Production + Box mips, block-row tiled over N GPUs
(strong scaling).  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port P profiles/tile_scaling.py [size] [steps] [p2p]
Every rank keeps a replica of the source image, builds the fp32 mip chain, encodes its block rows of every level and the
slices are gathered on rank 0 with NCCL (the only collective; it moves BCn bytes, 0.5 B/px) - or, with the `p2p` argument,
every rank's encode kernels store their blocks straight into rank 0's chain buffer over NVLink (CUDA IPC peer memory,
NvttbProcessDesc.bandOutputInPlace) and no collective follows the encode, only a barrier."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import nvtt_b200_loader  # noqa: E402

m = nvtt_b200_loader.load()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
p2p = len(sys.argv) > 3 and sys.argv[3] == "p2p"
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = m.Context(local)
img = m.synth.photo_bgra8(size, size, seed=1234)
adv = m.synth.adversarial_bgra8(size // 4, size // 4, seed=5)
img[: size // 4, : size // 4] = adv  # S1 + S5 mix
d_img = torch.from_numpy(img).to(dev)
desc = m.make_process_desc(0, size, size, m.Format_BC1, m.Quality_Production, mip_filter=0, band_index=rank, band_count=world)
n = int(m.lib().nvttb_process_output_size(desc))
sizes = [int(m.lib().nvttb_process_output_size(m.make_process_desc(0, size, size, m.Format_BC1, 2, mip_filter=0, band_index=b, band_count=world)))
         for b in range(world)]
cap = max(sizes)
mine = torch.zeros(cap, dtype=torch.uint8, device=dev)
gathered = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if world > 1 else None


shared = None
if p2p and world > 1:
    whole_n = int(m.lib().nvttb_process_whole_output_size(desc))
    shared = m.sharding.SharedOutput(ctx, whole_n, owner=0)
    desc_p2p = m.make_process_desc(0, size, size, m.Format_BC1, m.Quality_Production, mip_filter=0, band_index=rank, band_count=world,
                                   band_output_in_place=True)


def step():
    if shared is not None:
        ctx.process_to_device([d_img.data_ptr()], desc_p2p, shared.ptr, whole_n)
        ctx.synchronize()
        dist.barrier()  # every band's stores have landed in rank 0's buffer
        return
    ctx.process_to_device([d_img.data_ptr()], desc, mine.data_ptr(), cap)
    ctx.synchronize()
    if world > 1:
        dist.all_gather(gathered, mine)


for _ in range(2):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ok = None
if rank == 0:
    parts = [g[:s].cpu().numpy() for g, s in zip(gathered, sizes)] if world > 1 else [mine[:n].cpu().numpy()]
    d0 = m.make_process_desc(0, size, size, m.Format_BC1, 2, mip_filter=0, band_index=0, band_count=world)
    whole_desc = m.make_process_desc(0, size, size, m.Format_BC1, 2, mip_filter=0)
    nw = int(m.lib().nvttb_process_output_size(whole_desc))
    whole = torch.zeros(nw, dtype=torch.uint8, device=dev)
    ctx.process_to_device([d_img.data_ptr()], whole_desc, whole.data_ptr(), nw)
    ctx.synchronize()
    if shared is not None:
        # rank 0 owns the buffer (nvttb_device_alloc): read it back through the runtime API
        import cuda.bindings.runtime as rt
        host = np.empty(whole_n, np.uint8)
        torch.cuda.synchronize()
        (err,) = rt.cudaMemcpy(host.ctypes.data, shared.ptr, whole_n, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        assert int(err) == 0, err
        got = host
    else:
        got = m.sharding.assemble_bands(m.sharding.band_layout(m.lib(), d0, world), parts) if world > 1 else parts[0]
    ok = bool(np.array_equal(got, whole.cpu().numpy()))
    print(json.dumps({"workload": "configs[2]: %dx%d BGRA8 -> BC1 Production + Box mips, block-row tiled" % (size, size), "n_gpus": world,
                      "ms_per_image": float(ms.item()), "mpix_per_s": size * size / 1e6 / (float(ms.item()) / 1e3), "scaling": "strong",
                      "identical_to_single_gpu": ok, "collective": ("none: peer stores into rank 0's buffer over NVLink (CUDA IPC) + barrier" if shared is not None else "NCCL all_gather of BCn slices") if world > 1 else None}))
if shared is not None:
    dist.barrier()
    shared.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
