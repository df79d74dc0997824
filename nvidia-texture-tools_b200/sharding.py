"""Host-side sharding helpers for the multi-GPU path (one process per GPU, torch.distributed for the plumbing).
The encode itself never communicates: textures / cube faces are independent (src/nvtt/Context.cpp:260 loops over them
serially).  Only the owner of the OutputHandler gathers the BCn bytes of the other ranks."""
import numpy as np


def face_range(face_count, rank, world):
    """Contiguous, balanced split of faces/textures [0, face_count) over ranks; the first `rem` ranks get one extra."""
    base, rem = divmod(face_count, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_bytes(mine, dst=0):
    """Gathers variable-length uint8 arrays to rank `dst` in rank order (face-major order is preserved because
    face_range is contiguous).  Works with the gloo (CPU tensors) and nccl (CUDA tensors) backends."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = torch.tensor([mine.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes) if sizes else 0
    buf = torch.zeros(max(cap, 1), dtype=torch.uint8, device=dev)
    if mine.size:
        buf[:mine.size] = torch.from_numpy(np.ascontiguousarray(mine)).to(dev)
    out = [torch.zeros(max(cap, 1), dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(out, buf)
    if rank != dst:
        return None
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)]) if sizes else np.zeros(0, np.uint8)


def band_layout(lib, desc, world):
    """Per mip level and band: (offset, bytes, pitch, count) of the band's slices inside the level
    (nvttb_process_band_slices), for block-row sharding of ONE image over `world` GPUs (desc.bandChunkRows selects
    contiguous bands or cyclic chunks)."""
    import ctypes as C
    mips = lib.nvttb_process_mip_count(C.byref(desc))
    out = []
    for m in range(mips):
        row = []
        for b in range(world):
            d = type(desc).from_buffer_copy(desc)
            d.bandIndex, d.bandCount = b, world
            off, n, pitch, cnt = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_int(0)
            rc = lib.nvttb_process_band_slices(C.byref(d), m, C.byref(off), C.byref(n), C.byref(pitch), C.byref(cnt))
            assert rc == 0, rc
            row.append((off.value, n.value, pitch.value, cnt.value))
        out.append(row)
    return out


def assemble_bands(layout, per_band_bytes):
    """per_band_bytes[b] = the bytes band b produced (its slices of every level, concatenated).  Returns the whole
    mip chain (levels in order) exactly as a single GPU would emit it."""
    cur = [0] * len(per_band_bytes)
    levels = []
    for row in layout:
        size = max((off + (cnt - 1) * pitch + n) for off, n, pitch, cnt in row if cnt)
        lvl = np.zeros(size, np.uint8)
        for b, (off, n, pitch, cnt) in enumerate(row):
            for j in range(cnt):
                lvl[off + j * pitch:off + j * pitch + n] = per_band_bytes[b][cur[b]:cur[b] + n]
                cur[b] += n
        levels.append(lvl)
    return np.concatenate(levels)


def auto_chunk_rows(height, world):
    """Level-0 rows per chunk for cyclic block-row sharding: 4 * 2^j, whole chunks, the same number per band, about four
    chunks per band (mirrors auto_chunk_rows in csrc/capi.cu).  0: the height cannot be dealt out evenly."""
    best, c = 0, 4
    while c <= height:
        if height % c == 0 and (height // c) % world == 0:
            best = c
            if height // c // world <= 4:
                break
        c <<= 1
    return best


class SharedOutput:
    """ONE device buffer on rank `owner` that every rank's encoder writes its block rows into directly (peer memory over
    NVLink / NVSwitch): the gather of the BCn slices disappears into the encode kernels' stores.  The owner allocates with
    nvttb_device_alloc and exports a CUDA IPC handle; the 64 handle bytes go through torch.distributed; the others open it.
    `ptr` is the device pointer valid in this process (pass it to process_to_device with band_output_in_place=True)."""

    def __init__(self, ctx, nbytes, owner=0):
        import ctypes as C
        import torch
        import torch.distributed as dist
        self.ctx, self.owner, self.nbytes = ctx, owner, nbytes
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        L = ctx.L
        dev = torch.device("cuda", torch.cuda.current_device()) if (dist.is_initialized() and dist.get_backend() == "nccl") else torch.device("cpu")
        handle = torch.zeros(64, dtype=torch.uint8)
        p = C.c_void_p()
        if self.rank == owner:
            ctx._ck(L.nvttb_device_alloc(ctx.h, nbytes, C.byref(p)))
            buf = C.create_string_buffer(64)
            ctx._ck(L.nvttb_ipc_export(ctx.h, p, buf))
            handle = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if dist.is_initialized() and dist.get_world_size() > 1:
            h = handle.to(dev)
            dist.broadcast(h, src=owner)
            handle = h.cpu()
        if self.rank != owner:
            ctx._ck(L.nvttb_ipc_open(ctx.h, bytes(handle.numpy().tobytes()), C.byref(p)))
        self.ptr = p.value

    def close(self):
        import ctypes as C
        if self.ptr is None:
            return
        if self.rank == self.owner:
            self.ctx.L.nvttb_device_free(self.ctx.h, C.c_void_p(self.ptr))
        else:
            self.ctx.L.nvttb_ipc_close(self.ctx.h, C.c_void_p(self.ptr))
        self.ptr = None
