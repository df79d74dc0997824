"""Multi-GPU host logic on CPU: world_size-2 gloo run of the texture/face sharding + gather used by the N>1 path.
The data path itself has no collective (each rank encodes its own faces); only the final gather of BCn bytes to the
rank that owns the OutputHandler uses torch.distributed."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import nvtt_b200_loader
    m = nvtt_b200_loader.load()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    faces, w, h = 6, 16, 16
    lo, hi = m.sharding.face_range(faces, rank, world)
    # stand-in for the per-rank GPU encode (no GPU here): the oracle encodes this rank's faces
    import oracleapi
    imgs = [m.synth.photo_bgra8(w, h, seed=100 + f, alpha=True) for f in range(faces)]
    mine = oracleapi.process(imgs[lo:hi], 0, w, h, 4, 1, mip_filter=0) if hi > lo else np.zeros(0, np.uint8)
    whole = m.sharding.gather_bytes(mine, dst=0)
    if rank == 0:
        want = oracleapi.process(imgs, 0, w, h, 4, 1, mip_filter=0)
        q.put(bool(np.array_equal(whole, want)))
    dist.barrier()
    dist.destroy_process_group()


def _band_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import nvtt_b200_loader
    import oracleapi
    m = nvtt_b200_loader.load()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ok = True
    # contiguous bands (bandChunkRows = 0) and cyclic chunks of 8 level-0 rows (chunk c belongs to band c % world)
    for w, h, chunk in ((64, 32, 0), (64, 32, 8), (40, 48, 4)):
        img = m.synth.photo_bgra8(w, h, seed=9, alpha=True)
        d = m.make_process_desc(0, w, h, 1, 1, mip_filter=0, band_index=rank, band_count=world, band_chunk_rows=chunk)
        layout = m.sharding.band_layout(m.lib(), d, world)
        # stand-in for this rank's GPU encode of its block rows (no GPU here): the oracle's whole chain, cut by the contract
        whole = oracleapi.process([img], 0, w, h, 1, 1, mip_filter=0)
        mine, base = [], 0
        for row in layout:
            off, n, pitch, cnt = row[rank]
            for j in range(cnt):
                mine.append(whole[base + off + j * pitch:base + off + j * pitch + n])
            base += max(o + (c - 1) * p + k for o, k, p, c in row if c)
        mine = np.concatenate(mine) if mine else np.zeros(0, np.uint8)
        sizes_ok = mine.size == int(m.lib().nvttb_process_output_size(d))
        allb = m.sharding.gather_bytes(mine, dst=0)
        if rank == 0:
            per_band, p = [], 0
            for b in range(world):
                db = m.make_process_desc(0, w, h, 1, 1, mip_filter=0, band_index=b, band_count=world, band_chunk_rows=chunk)
                n = int(m.lib().nvttb_process_output_size(db))
                per_band.append(allb[p:p + n])
                p += n
            got = m.sharding.assemble_bands(layout, per_band)
            ok = ok and bool(sizes_ok and np.array_equal(got, whole))
            # every block row of every level belongs to exactly one band
            for row in layout:
                covered = sorted((o + j * pt, o + j * pt + k) for o, k, pt, c in row for j in range(c))
                ok = ok and covered[0][0] == 0 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    if rank == 0:
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_block_row_bands():
    """Block-row sharding of one image (configs[2]): layout contract + gather + re-assembly, world_size 2 on gloo."""
    import oracleapi
    if not oracleapi.available():
        pytest.skip("oracle not built")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_band_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok
    assert all(p.exitcode == 0 for p in procs)


def test_face_ranges_cover_everything(nvtt):
    for faces in (1, 6, 7, 64):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, hi = nvtt.sharding.face_range(faces, r, world)
                assert 0 <= lo <= hi <= faces
                got += list(range(lo, hi))
            assert got == list(range(faces))
            sizes = [nvtt.sharding.face_range(faces, r, world) for r in range(world)]
            assert max(b - a for a, b in sizes) - min(b - a for a, b in sizes) <= 1


def test_two_rank_gloo_gather():
    import oracleapi
    if not oracleapi.available():
        pytest.skip("oracle not built")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok
    assert all(p.exitcode == 0 for p in procs)
