#!/usr/bin/env python
"""bench.py — BCn encode throughput (Mpix/s incl. mips) of the B200-native path, per the driver contract.

Workload at every N (weak scaling, texture-sharded, no data-path collective): BASELINE.json configs[1] —
per GPU and per step one 4096x4096 BGRA8 colour+alpha texture -> BC3 (Quality_Normal) and one 4096x4096 normal
map -> BC5 (Quality_Normal), both with the full Kaiser(3,4,1) mip chain, through the whole InputOptions pipeline
(setImage -> toLinear -> [mips -> renormalise] -> toGamma -> block encode).

  value  : level-0 Mpix/s, inputs (BGRA8) and outputs (BCn) resident in HBM, CUDA events on the launching stream.
  e2e    : same metric through the C-ABI call with HOST buffers (pinned): H2D of the texels and D2H of the encoded
           chain happen inside the timed region, wall clock around a device synchronize.
  roofline / cpu_baseline: see DESIGN.md §Measurement.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, all host threads) instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bcn_encode_mpix_per_s_incl_mips"
UNIT = "Mpix/s"
SIZE = 4096
REF_SAMPLE = 1024  # the CPU arm runs this crop-size version of the same workload per step (bounded)


def chain_pixels(w, h):
    t = 0
    while True:
        t += w * h
        if w == 1 and h == 1:
            return t
        w, h = max(1, w // 2), max(1, h // 2)


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def run_reference(args, world, rank):
    """Reference arm: the reference's CPU encoders (fast build if present, else the pinned build), all host threads."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refapi
    import nvtt_b200_loader
    m = nvtt_b200_loader.load()
    fast = refapi.available(fast=True)
    if not fast and not refapi.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built"}))
        return
    n = REF_SAMPLE
    a = m.synth.photo_bgra8(n, n, seed=1234, alpha=True)
    nm = m.synth.normal_bgra8(n, n, seed=7)

    def step():
        refapi.process([a], 0, n, n, refapi.Format_BC3, 1, mip_filter=2, fast=fast)
        refapi.process([nm], 0, n, n, refapi.Format_BC5, 1, mip_filter=2, normal_map=True, fast=fast)

    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    mpix = 2 * n * n / 1e6
    v = mpix * args.steps / dt
    cores = os.cpu_count()
    sample = "%dx%d BC3(S2)+BC5(S3) Kaiser mip chains per step (same pipeline, 1/16 of the 4096^2 texels)" % (n, n)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                         "build": "fast (SSE2 icbc/squish, -O3)" if fast else "pinned (-O2 scalar)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {
        "workload": "configs[1]: BC3 colour+alpha (S2) + BC5 normal map (S3), 4096x4096 BGRA8 each, Kaiser(3,4,1) full mip "
                    "chain, Quality_Normal, wrap Mirror, gamma 2.2/2.2; one such pair per GPU per step",
        "pixels_level0_per_step_per_gpu": 2 * SIZE * SIZE,
        "pixels_incl_mips_per_step_per_gpu": 2 * chain_pixels(SIZE, SIZE),
        "mpix_definition": "level-0 pixels / time for the whole chain",
        "parallelism": "texture-sharded x%d (no collective)" % n_gpus,
        "l2": "inputs larger than L2 (64 MiB BGRA8 + 256 MiB planar fp32 level 0 per texture vs 126 MB L2)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world, rank, local = dist_setup()
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import nvtt_b200_loader
    m = nvtt_b200_loader.load()

    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ctx = m.Context(local)  # raises if the CUDA library / GPU is missing: there is no fallback

    # ---- synthetic inputs (seeded, per rank) --------------------------------------------------------------
    col = m.synth.photo_bgra8(SIZE, SIZE, seed=1234 + rank, alpha=True)
    nrm = m.synth.normal_bgra8(SIZE, SIZE, seed=7 + rank)
    h_col = torch.from_numpy(col).pin_memory()
    h_nrm = torch.from_numpy(nrm).pin_memory()
    d_col = h_col.to(dev)
    d_nrm = h_nrm.to(dev)
    desc3 = m.make_process_desc(m.InputFormat_BGRA_8UB, SIZE, SIZE, m.Format_BC3, m.Quality_Normal,
                                mip_filter=m.MipmapFilter_Kaiser, wrap=m.WrapMode_Mirror)
    desc5 = m.make_process_desc(m.InputFormat_BGRA_8UB, SIZE, SIZE, m.Format_BC5, m.Quality_Normal,
                                mip_filter=m.MipmapFilter_Kaiser, wrap=m.WrapMode_Mirror, normal_map=True)
    out_bytes = int(m.lib().nvttb_process_output_size(desc3))
    d_out3 = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
    d_out5 = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def step_device():
        ctx.process_to_device([d_col.data_ptr()], desc3, d_out3.data_ptr(), out_bytes)
        ctx.process_to_device([d_nrm.data_ptr()], desc5, d_out5.data_ptr(), out_bytes)

    import ctypes as C
    emitted = [0]

    def _emit(user, face, mip, w, h, d, data, size):
        emitted[0] += size  # the encoded chain is already in (pinned) host memory here; a real handler would write it out
        return 1

    emit_cb = m.capi.EMIT_FN(_emit)
    p_col = (C.c_void_p * 1)(h_col.data_ptr())
    p_nrm = (C.c_void_p * 1)(h_nrm.data_ptr())

    def step_e2e_sequential():
        ctx._ck(ctx.L.nvttb_process(ctx.h, C.byref(desc3), p_col, m.HOST, emit_cb, None))
        ctx._ck(ctx.L.nvttb_process(ctx.h, C.byref(desc5), p_nrm, m.HOST, emit_cb, None))

    # The two textures of a step are independent: a caller with a second context encodes them from two host threads, so
    # that one texture's copies run under the other's kernels (the C ABI is thread-safe per context; the reference would
    # serialise the two process() calls on its global thread pool).
    from concurrent.futures import ThreadPoolExecutor
    ctx2 = m.Context(local)
    pool = ThreadPoolExecutor(2)

    def _one(c, desc, ptrs):
        c._ck(c.L.nvttb_process(c.h, C.byref(desc), ptrs, m.HOST, emit_cb, None))

    def step_e2e():
        a = pool.submit(_one, ctx, desc3, p_col)
        b = pool.submit(_one, ctx2, desc5, p_nrm)
        a.result()
        b.result()

    def barrier():
        if use_dist:
            dist.barrier()
        ctx.synchronize()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput --------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    ctx.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = ctx.timer_stop()
    barrier()
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    mpix_step = 2 * SIZE * SIZE / 1e6
    value = world * mpix_step * args.steps / (ms / 1e3)

    # ---- end to end through the C ABI with host buffers --------------------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    ctx.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    dt = max_over_ranks(dt)
    e2e_value = world * mpix_step * args.steps / dt
    # the same step with one context and one host thread (the two calls back to back), for comparison
    for _ in range(2):
        step_e2e_sequential()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e_sequential()
    ctx.synchronize()
    dt_seq = time.perf_counter() - t0
    barrier()
    dt_seq = max_over_ranks(dt_seq)

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (CUDA events around every launch of one extra step) ---------------
        ctx.profile_begin()
        step_device()
        prof = ctx.profile_end()
        tot = sum(v["total_ms"] for v in prof.values())
        dom = max(prof.items(), key=lambda kv: kv[1]["total_ms"])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # algorithmic bytes per level-0 pixel for BC3/BC5 from BGRA8: 4 + 1*4/3 = 5.333; one launch of the
        # dominant kernel covers max_units texels of one level.
        bpp = 4.0 + 1.0  # bytes per texel of the level the launch covers (source texel in, 1 B/px BCn out)
        k = dom[1]
        achieved = bpp * k["max_units"] / (k["max_ms"] / 1e3) / 1e9
        sm_mhz = (clocks or {}).get("sm_mhz") or 0.0
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "kernel": dom[0], "kernel_ms_per_launch": k["max_ms"], "kernel_share_of_step": k["total_ms"] / tot,
            "peak_source": peak_src,
            "note": "the dominant kernel is FP32-issue bound, not HBM bound (no dense contraction, tensor cores unused); "
                    "issue-slot utilisation from ncu is in profiles/",
            "per_kernel_ms": {n: round(v["total_ms"], 4) for n, v in sorted(prof.items())},
        }
        if dom[0] == "k_bc3_color":
            # From the committed ncu --set full capture of this kernel on this input (profiles/r1i_ncu_bc3_color.txt):
            # DRAM bytes and warp instructions of the level-0 launch.  The live CUDA-event time of the same launch turns the
            # instruction count into an issue rate against 148 SMs x 4 schedulers x 1 warp-instruction per clock.
            # dram__bytes_read 50.4 MB + dram__bytes_write 0.06 MB at 2048x2048 (the 4 MB of blocks were still in L2 when the capture
            # ended), scaled by the texel count of this launch; algorithmic bytes are 5 B/texel = 83.9 MB at 4096x4096
            roofline["traffic"] = (50.4e6 + 0.06e6) * (SIZE * SIZE) / (2048.0 * 2048.0)
            winst = 9505.0 * (SIZE // 4) * (SIZE // 4) / 2  # 9505 warp-instructions per warp of two 4x4 blocks
            if sm_mhz:
                issue_peak = 148 * 4 * sm_mhz * 1e6
                issue_ach = winst / (k["max_ms"] / 1e3)
                roofline["issue"] = {"bound": "sm_issue", "achieved": issue_ach, "peak": issue_peak, "unit": "warp-inst/s",
                                     "frac": issue_ach / issue_peak, "warp_insts_per_launch": winst,
                                     "fma_pipe_active_pct_ncu": 72.7,
                                     "source": "smsp__inst_executed.sum per warp of the level-0 launch (profiles/r1i_ncu_bc3_color.txt, same synthetic image "
                                               "at 2048x2048) x the warps of this launch; the kernel is FP32-pipe bound (math-pipe throttle is its top stall); "
                                               "time = this run's CUDA events; clock = median SM clock sampled during the timed region"}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_baseline = cpu_baseline_leg(m, ctx)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * SIZE * SIZE * 4,
                    "d2h_bytes_per_step": 2 * out_bytes, "ms_per_step": dt / args.steps * 1e3,
                    "how": "nvttb_process with pinned host buffers; the step's two textures on two contexts from two host threads",
                    "one_context_sequential": {"value": world * mpix_step * args.steps / dt_seq, "ms_per_step": dt_seq / args.steps * 1e3}},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
    barrier()
    if use_dist:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def cpu_baseline_leg(m, ctx=None):
    """The reference's CPU path on the host cores of this box, on a bounded sample of the same workload.  Its output is
    also the checker of the metric's second half ("PSNR delta vs ref CPU"): our encode of the same sample is compared
    block by block, and both level-0 images are decoded and measured on the GPU (nvttb_surface_set_image_2d + rmsError)."""
    import math
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import refapi
    except Exception as e:  # pragma: no cover
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "unavailable: %s" % e}
    fast = refapi.available(fast=True)
    if not fast and not refapi.available():
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "oracle/_ref not built"}
    n = 2048
    a = m.synth.photo_bgra8(n, n, seed=1234, alpha=True)
    nm = m.synth.normal_bgra8(n, n, seed=7)
    t0 = time.perf_counter()
    ref3 = refapi.process([a], 0, n, n, refapi.Format_BC3, 1, mip_filter=2, fast=fast)
    ref5 = refapi.process([nm], 0, n, n, refapi.Format_BC5, 1, mip_filter=2, normal_map=True, fast=fast)
    dt = time.perf_counter() - t0
    out = {"value": 2 * n * n / 1e6 / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
           "sample": "one %dx%d BC3(S2)+BC5(S3) Kaiser mip-chain pair (1/4 of the step's texels), %s build, %.1f s"
                     % (n, n, "fast SSE2 -O3" if fast else "pinned -O2 scalar", dt)}
    if ctx is not None:
        try:
            d3 = m.make_process_desc(m.InputFormat_BGRA_8UB, n, n, m.Format_BC3, m.Quality_Normal, mip_filter=m.MipmapFilter_Kaiser)
            d5 = m.make_process_desc(m.InputFormat_BGRA_8UB, n, n, m.Format_BC5, m.Quality_Normal, mip_filter=m.MipmapFilter_Kaiser, normal_map=True)
            our3, our5 = ctx.process_bytes([a], d3), ctx.process_bytes([nm], d5)
            pinned = refapi.available()
            if pinned and fast:  # the checker is the pinned scalar build (untimed); the fast build above is only the timing arm
                ref3 = refapi.process([a], 0, n, n, refapi.Format_BC3, 1, mip_filter=2)
                ref5 = refapi.process([nm], 0, n, n, refapi.Format_BC5, 1, mip_filter=2, normal_map=True)
            par = {}
            for name, fmt, src, ours, theirs in (("bc3", m.Format_BC3, a, our3, ref3), ("bc5", m.Format_BC5, nm, our5, ref5)):
                same = ours.size == theirs.size
                bad = int((ours.reshape(-1, 16) != theirs.reshape(-1, 16)).any(1).sum()) if same else -1
                org = m.Surface(ctx)
                if name == "bc5":  # BC5 stores x, y only (decoded blue is 0): measure those two channels
                    src = src.copy()
                    src[..., 0] = 0
                org.set_image(m.InputFormat_BGRA_8UB, n, n, src)
                lvl0 = (n // 4) * (n // 4) * 16
                psnr = []
                for data in (ours, theirs):
                    dec = m.Surface(ctx)
                    dec.set_image_2d(fmt, n, n, data[:lvl0])
                    rms = org.rms_error(dec)
                    psnr.append(20.0 * math.log10(1.0 / rms) if rms > 0 else float("inf"))
                par[name] = {"blocks": int(theirs.size // 16), "mismatching_blocks": bad, "psnr_level0_db": psnr[0],  # 20 log10(1 / nvtt::rmsError), rmsError = sqrt(sum over r,g,b / texels)
                             "psnr_level0_reference_db": psnr[1], "psnr_delta_db": psnr[0] - psnr[1]}
            out["parity_vs_reference_output"] = par
            out["parity_note"] = "checker: %s reference build on the same %dx%d sample" % ("pinned scalar" if pinned else "fast SSE", n, n)
        except Exception as e:  # pragma: no cover
            out["parity_vs_reference_output"] = "failed: %s" % e
    return out


if __name__ == "__main__":
    main()
