"""GPU tests of the multi-GPU paths of ONE process_*/shard call chain, exercised on one GPU with several contexts (each
context stands for one GPU of the box: own streams, own buffers; the exchange buffer and the shared output are plain
device memory here and peer memory on a real box).  Bar: the assembled chain is byte-identical to the single-GPU chain
and, through it, to the reference."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _contexts(nvtt, n, spread=False):
    """n contexts on GPU 0 - or, with spread, dealt over the GPUs of the box (one GPU: all on it)."""
    ndev = max(1, nvtt.lib().nvttb_device_count())
    return [nvtt.Context(i % ndev if spread else 0) for i in range(n)]


def _pinned(a):
    """Host buffers of bands that share one GPU must be pinned (pageable copies are synchronous in the driver and wait while
    band 0's device-side wait is resident).  Returns (numpy view, owner)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t


CASES = [
    # w, h, format, quality, world, chunk rows, extra
    (256, 256, "BC1", 2, 4, 16, {}),
    (256, 256, "BC1", 2, 2, 32, {}),
    (512, 256, "BC3", 1, 4, 8, {}),
    (128, 512, "BC5", 1, 8, 16, dict(normal_map=True)),
    (256, 256, "BC4", 1, 2, 0, {}),          # contiguous bands
    (64, 64, "BC7", 1, 2, 8, {}),
    (128, 128, "BC1", 1, 4, 4, dict(gamma=(1.0, 1.0))),
    (256, 512, "BC1", 2, 2, 64, {}),         # chunks of >= 64 rows: host input uploads / encodes the first chunk in two pieces
    (128, 1024, "BC3", 1, 4, 128, {}),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d-%s-q%d-n%d-c%d" % c[:6])
def test_cyclic_chunks_replicated_front_end(nvtt, ref, ctx, case):
    """bandChunkRows: chunk c of level 0 belongs to band c % N; every band builds the whole chain and encodes its chunks."""
    w, h, fmt_name, q, world, chunk, kw = case
    fmt = getattr(nvtt, "Format_" + fmt_name)
    img = nvtt.synth.photo_bgra8(w, h, seed=3, alpha=True) if not kw.get("normal_map") else nvtt.synth.normal_bgra8(w, h, seed=3)
    whole = ctx.process_bytes([img], nvtt.make_process_desc(0, w, h, fmt, q, **kw))
    if fmt_name != "BC7":
        assert np.array_equal(whole, ref.process([img], 0, w, h, fmt, q, **kw))
    parts = [ctx.process_bytes([img], nvtt.make_process_desc(0, w, h, fmt, q, band_index=b, band_count=world, band_chunk_rows=chunk, **kw))
             for b in range(world)]
    d0 = nvtt.make_process_desc(0, w, h, fmt, q, band_index=0, band_count=world, band_chunk_rows=chunk, **kw)
    got = nvtt.sharding.assemble_bands(nvtt.sharding.band_layout(nvtt.lib(), d0, world), parts)
    assert np.array_equal(got, whole)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d-%s-q%d-n%d-c%d" % c[:6])
@pytest.mark.parametrize("host_input", [False, True])
def test_band_local_front_end_with_exchange(nvtt, ctx, case, host_input):
    """Each band uploads / converts / down-samples only its own chunks; the last distributed level reaches band 0 through
    the exchange buffer (flags with release / acquire semantics) and band 0 runs the tail on its second stream.  Three
    images in a row through the same exchange buffer (sequence numbers, double-buffered rows, acknowledge word)."""
    import torch
    w, h, fmt_name, q, world, chunk, kw = case
    fmt = getattr(nvtt, "Format_" + fmt_name)
    L = nvtt.lib()
    d0 = nvtt.make_process_desc(0, w, h, fmt, q, band_index=0, band_count=world, band_chunk_rows=chunk, **kw)
    xbytes = int(L.nvttb_process_exchange_size(C.byref(d0)))
    assert xbytes > 0
    nw = int(L.nvttb_process_whole_output_size(C.byref(d0)))
    xchg = C.c_void_p()
    ctx._ck(L.nvttb_device_alloc(ctx.h, xbytes, C.byref(xchg)))
    ctxs = _contexts(nvtt, world)
    shared = torch.zeros(nw, dtype=torch.uint8, device="cuda")
    try:
        for b in range(world):  # the bands share ONE GPU here: no allocation while band 0 waits on the device
            ctxs[b].process_prepare(nvtt.make_process_desc(0, w, h, fmt, q, band_index=b, band_count=world, band_chunk_rows=chunk,
                                                           band_output_in_place=True, band_exchange=xchg.value, band_sequence=1, **kw),
                                    location=nvtt.HOST if host_input else nvtt.DEVICE, own_output=False)
        for seq in (1, 2, 3, 4):
            img = (nvtt.synth.photo_bgra8(w, h, seed=10 + seq, alpha=True) if not kw.get("normal_map")
                   else nvtt.synth.normal_bgra8(w, h, seed=10 + seq))
            whole = ctx.process_bytes([img], nvtt.make_process_desc(0, w, h, fmt, q, **kw))
            shared.zero_()
            torch.cuda.synchronize()
            d_img = torch.from_numpy(img).cuda()
            pimg, _keep = _pinned(img)
            # band 0 last: with host input the call blocks until the texels are consumed, and band 0's tail waits for the others
            for b in list(range(1, world)) + [0]:
                d = nvtt.make_process_desc(0, w, h, fmt, q, band_index=b, band_count=world, band_chunk_rows=chunk,
                                           band_output_in_place=True, band_exchange=xchg.value, band_sequence=seq, **kw)
                if host_input:
                    ctxs[b].process_to_device([pimg], d, shared.data_ptr(), nw, location=nvtt.HOST)
                else:
                    ctxs[b].process_to_device([d_img.data_ptr()], d, shared.data_ptr(), nw)
            for c in ctxs:
                c.synchronize()
            assert np.array_equal(shared.cpu().numpy(), whole), "image %d" % seq
    finally:
        for c in ctxs:
            c.close()
        L.nvttb_device_free(ctx.h, xchg)


def test_shard_to_host_buffer(nvtt, ctx):
    """nvttb_process_shard with a host output: every band copies its slices into ONE host buffer in the whole-chain layout."""
    w, h, world, chunk = 256, 256, 4, 16
    L = nvtt.lib()
    img, _keep_img = _pinned(nvtt.synth.photo_bgra8(w, h, seed=21, alpha=True))
    for fmt, q, kw in ((nvtt.Format_BC1, 2, {}), (nvtt.Format_BC3, 1, dict(mip_filter=2))):  # Kaiser: replicated front end
        whole = ctx.process_bytes([img], nvtt.make_process_desc(0, w, h, fmt, q, **kw))
        d0 = nvtt.make_process_desc(0, w, h, fmt, q, band_index=0, band_count=world, band_chunk_rows=chunk, **kw)
        xbytes = int(L.nvttb_process_exchange_size(C.byref(d0)))
        xchg = C.c_void_p()
        if xbytes:
            ctx._ck(L.nvttb_device_alloc(ctx.h, xbytes, C.byref(xchg)))
        host, _keep_host = _pinned(np.zeros(whole.size, np.uint8))
        ctxs = _contexts(nvtt, world)
        try:
            import threading
            errs = []
            for b in range(world):  # the bands share ONE GPU here: no allocation while band 0 waits on the device
                ctxs[b].process_prepare(nvtt.make_process_desc(0, w, h, fmt, q, band_index=b, band_count=world, band_chunk_rows=chunk,
                                                               band_output_in_place=True, band_exchange=xchg.value, band_sequence=1, **kw))

            def run(b):
                try:
                    d = nvtt.make_process_desc(0, w, h, fmt, q, band_index=b, band_count=world, band_chunk_rows=chunk,
                                               band_output_in_place=True, band_exchange=xchg.value, band_sequence=1, **kw)
                    ctxs[b].process_shard([img], d, None, host.ctypes.data)
                except Exception as e:  # pragma: no cover
                    errs.append(e)

            th = [threading.Thread(target=run, args=(b,)) for b in range(world)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            assert not errs, errs
            assert np.array_equal(host, whole)
        finally:
            for c in ctxs:
                c.close()
            if xbytes:
                L.nvttb_device_free(ctx.h, xchg)


def test_process_multi(nvtt, ref, ctx):
    """nvttb_process_multi: one call, several contexts (GPUs), host images in, the single-GPU emit sequence out."""
    ctxs = _contexts(nvtt, 4, spread=True)  # on a multi-GPU box: really one context per GPU (band-local front end, peer exchange)
    try:
        # one large image: block-row sharded (Box: band-local front end; Kaiser: replicated front end)
        img = nvtt.synth.photo_bgra8(1024, 1024, seed=2, alpha=True)
        for fmt, q, kw in ((nvtt.Format_BC1, 1, {}), (nvtt.Format_BC3, 1, dict(mip_filter=2)), (nvtt.Format_BC1, 1, dict(mipmaps=False))):
            d = nvtt.make_process_desc(0, 1024, 1024, fmt, q, **kw)
            want = ctx.process([img], d)
            for _ in range(3):  # consecutive images through the same exchange buffer
                got = nvtt.capi.process_multi(ctxs, [img], d)
                assert [(f, m, w, h) for f, m, w, h, _ in got] == [(f, m, w, h) for f, m, w, h, _ in want]
                assert all(np.array_equal(a[4], b[4]) for a, b in zip(got, want))
        # cube faces: dealt out face by face
        faces = [nvtt.synth.hdr_rgba16f(64, 64, seed=30 + i) for i in range(6)]
        d = nvtt.make_process_desc(nvtt.InputFormat_RGBA_16F, 64, 64, nvtt.Format_BC6, 1, faces=6, pixel_type=nvtt.PixelType_UnsignedFloat)
        want = ctx.process(faces, d)
        got = nvtt.capi.process_multi(ctxs, faces, d)
        assert len(got) == len(want) and all(a[:4] == b[:4] and np.array_equal(a[4], b[4]) for a, b in zip(got, want))
        # small image, one face: falls through to the first context
        small = nvtt.synth.photo_bgra8(100, 60, seed=4)
        d = nvtt.make_process_desc(0, 100, 60, nvtt.Format_BC1, 1)
        assert all(np.array_equal(a[4], b[4]) for a, b in zip(nvtt.capi.process_multi(ctxs, [small], d), ctx.process([small], d)))
        assert np.array_equal(np.concatenate([b[4] for b in ctx.process([small], d)]), ref.process([small], 0, 100, 60, nvtt.Format_BC1, 1))
    finally:
        for c in ctxs:
            c.close()


def test_emit_path_rejects_in_place_bands(nvtt, ctx):
    img = nvtt.synth.photo_bgra8(64, 64, seed=1)
    d = nvtt.make_process_desc(0, 64, 64, nvtt.Format_BC1, 1, band_index=1, band_count=2, band_output_in_place=True)
    with pytest.raises(nvtt.capi.NvttbError) as e:
        ctx.process([img], d)
    assert e.value.code == 2  # Error_InvalidInput
