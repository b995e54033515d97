#!/usr/bin/env python
"""bench.py -- Gpixel/s of fused forward DCT+quantise and dequantise+IDCT on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 3840x2160 8-bit grayscale frames, quality 50, adaptive off,
NATURAL layout.  One STEP = one pass of the hot path over one batch of `--frames` synthetic frames:
K1 (u8 pixels -> int16 records) followed by K2 (records -> u8 pixels), both through libdct_cuda's
C ABI.  The batch (64 frames = 531 MB of pixels + 1.06 GB of records) is larger than the 126 MB
L2, so no flush is needed between iterations.  `value` counts every pixel once per direction:
    value = 2 * frames * W * H * n_gpus / step_time           [Gpixel/s, device-resident]
`e2e` is the same metric through the host-plane calls with PINNED HOST buffers, the H2D and D2H
copies inside the timed region; next to it `e2e.copy_ceiling_gbs` is what bare cudaMemcpyAsync
reaches on the same box, same byte counts, all ranks at once (H2D only, D2H only, both directions
at once) and `e2e.frac_of_ceiling` the step's traffic against the duplex figure.  `roofline` is
for K1 (the dominant kernel): 192 algorithmic bytes per 8x8 block (64 B of pixels in, 128 B of
records out) over the K1 launch's own CUDA-event time, against the measured copy
bandwidth in MEASURED_PEAKS.json; `roofline.step_frac` is the same for the whole step and `kernels`
lists the four launches of a step.  `parity_ok`: every rank checks a strip of ITS output (forward
records, inverse pixels, tie count) against the CPU oracle; MIN over ranks.  `shapes` carries every
BASELINE shape (C1..C5), device-resident and end to end through the host-plane calls.
`cpu_baseline` times the reference's own C code (oracle/_ref, else the oracle port) on this box.

Multi-GPU: every rank owns its own batch (weak scaling), no data-path collective; torch.distributed
(NCCL) is used only for the barrier and the max-over-ranks of the device time.
`--impl reference` times the reference's CPU implementation on all host threads instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 3840, 2160
QUALITY = 50
METRIC = "gpixel_per_s_fwd_dct_quant_plus_dequant_idct"
UNIT = "Gpixel/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="4K frames per batch (per GPU)")
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the CPU baseline sample (0 = auto)")
    ap.add_argument("--layout", type=int, default=0)
    ap.add_argument("--adaptive", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shapes", action="store_true", help="skip the per-shape block (C1..C5)")
    ap.add_argument("--quality", type=int, default=QUALITY, help="non-default values are exploratory, not the headline config")
    return ap.parse_args()


def workload_config(args):
    """The SAME dict in both arms (the driver compares them key by key)."""
    return {"workload": f"{W}x{H} 8-bit grayscale frames, q{args.quality}, adaptive {'on' if args.adaptive else 'off'}: "
                        f"forward DCT+quantise then dequantise+IDCT (BASELINE configs[1])",
            "quality": args.quality, "adaptive": int(args.adaptive), "layout": "zigzag" if args.layout else "natural",
            "l2": "inputs larger than L2 (a step reads/writes > 1 GB per direction), no flush",
            "sharding": "one batch per GPU, no collective on the data path"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """SM clock, power and throttle reasons sampled through NVML DURING the timed region
    (a background thread, one sample every ~2 ms; nvidia-smi's own loop is too slow to start)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, torch_index):
        self.samples, self.stop_flag, self.thread, self.h, self.err = [], False, None, None, None
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def pcie(self):
        """PCIe link of this GPU as NVML reports it (generation x width, current and maximum)."""
        if self.h is None:
            return None
        nv = self.nv
        try:
            return {"gen": nv.nvmlDeviceGetCurrPcieLinkGeneration(self.h), "width": nv.nvmlDeviceGetCurrPcieLinkWidth(self.h),
                    "gen_max": nv.nvmlDeviceGetMaxPcieLinkGeneration(self.h), "width_max": nv.nvmlDeviceGetMaxPcieLinkWidth(self.h)}
        except Exception as e:  # pragma: no cover
            return {"error": repr(e)}

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, pw, rs))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.h is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err or "no samples"}
        inside = [x for x in self.samples if t0 <= x[0] <= t1] or self.samples[-3:]
        mask = 0
        for x in inside:
            mask |= int(x[3])
        return {"sm_mhz": statistics.median(x[1] for x in inside), "sm_max_mhz": self.max_sm,
                "power_w_max": max(x[2] for x in inside), "samples": len(inside),
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b), "how": "NVML, ~2 ms period, inside the timed region"}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own C code on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_checker(which="ref"):
    from oracle import binding
    if which == "ref" and binding.have_ref():
        return binding.load("ref"), "reference"
    if which == "ref_shipped":
        try:
            return binding.load("ref_shipped"), "reference"
        except Exception:
            return None, None
    return binding.load("oracle"), "port"


def cpu_run(frames, threads, quality=QUALITY, which="ref", rows=None):
    """fwd+inv of `frames` 4K frames (or `rows` pixel rows) through the reference's block functions; (Gpx/s, s, kind)."""
    import numpy as np
    lib, kind = cpu_checker(which)
    if lib is None:
        return None, None, None
    rng = np.random.default_rng(1)
    nrows = rows if rows is not None else frames * H
    px = rng.integers(0, 256, size=(nrows, W), dtype=np.uint8)
    Q = lib.quant_table(quality)
    t0 = time.perf_counter()
    coef, _, _ = lib.fwd_quant_plane(px, Q, 0, 0, nthreads=threads)
    lib.dequant_idct_plane(coef, W, nrows, Q, 0, 0, None, nthreads=threads)
    dt = time.perf_counter() - t0
    return 2.0 * nrows * W / dt / 1e9, dt, kind


def reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = args.cpu_frames or max(2, min(16, threads // 2))
    cpu_run(1, threads, args.quality)                      # warm the page cache / thread pool once
    times = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        v, dt, kind = cpu_run(frames, threads, args.quality)
        if i >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    value = 2.0 * frames * W * H / t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "batch": {"frames_per_step": frames, "host_threads": threads,
                  "note": "a bounded sample of the same workload per step; throughput of a CPU arm does not depend on the batch size"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{frames} frames of {W}x{H} per step, fwd+inv, {threads} pthreads over block rows, -O2 -std=c99"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm helpers
# ---------------------------------------------------------------------------------------------
class Ctx:
    """rank / device / barrier plumbing shared by the legs of the GPU arm"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return float(t.item())

    def wall(self, fn):
        """host wall clock around fn, barrier + synchronize on both sides, max over ranks"""
        self.barrier()
        t0 = time.perf_counter()
        fn()
        self.barrier()
        return self.max_over_ranks(time.perf_counter() - t0)


def copy_ceiling(cx, h2d_bytes, d2h_bytes, reps=4):
    """Bare pinned-memory cudaMemcpyAsync on this box, all ranks at once: what the PCIe / host side can deliver
    for the byte counts of one e2e step.  GB/s per direction, whole job (all ranks), slowest rank's time."""
    torch = cx.torch
    h_src = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_dst = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    h_src.fill_(3)
    d_dst = torch.empty(h2d_bytes, dtype=torch.uint8, device=cx.dev)
    d_src = torch.full((d2h_bytes,), 5, dtype=torch.uint8, device=cx.dev)
    s_up, s_dn = torch.cuda.Stream(device=cx.dev), torch.cuda.Stream(device=cx.dev)

    def up():
        with torch.cuda.stream(s_up):
            for _ in range(reps):
                d_dst.copy_(h_src, non_blocking=True)

    def down():
        with torch.cuda.stream(s_dn):
            for _ in range(reps):
                h_dst.copy_(d_src, non_blocking=True)

    def both():
        up()
        down()

    out = {}
    for name, fn, nbytes in (("h2d", up, h2d_bytes), ("d2h", down, d2h_bytes), ("duplex", both, None)):
        fn()
        cx.barrier()
        t = min(cx.wall(fn) for _ in range(2)) / reps
        if nbytes is None:
            out["duplex_h2d"] = h2d_bytes * cx.world / t / 1e9
            out["duplex_d2h"] = d2h_bytes * cx.world / t / 1e9
            out["duplex_ms_per_step"] = t * 1e3
        else:
            out[name] = nbytes * cx.world / t / 1e9
    out["what"] = (f"bare cudaMemcpyAsync, pinned host memory, {h2d_bytes / 1e6:.0f} MB up / {d2h_bytes / 1e6:.0f} MB down per rank "
                   f"per step, {cx.world} rank(s) at once; GB/s over all ranks, slowest rank's time")
    del h_src, h_dst, d_dst, d_src
    return out


def parity_check(cx, plan, px, coef, rec, layout, quality, adaptive, frames):
    """One strip of THIS rank's batch against the CPU oracle: forward records, inverse pixels, tie count.
    The strip (64 pixel rows of a frame in the middle of the batch) goes through the same device calls as the
    timed step; the whole batch's records / pixels that the step left behind are compared on that strip too."""
    import numpy as np

    from dct_b200 import api
    from oracle import binding
    torch = cx.torch
    orc = binding.load("oracle")
    r0 = (frames // 2) * H + 1024          # a multiple of 8
    nrows = 64
    strip = px[r0:r0 + nrows].contiguous()
    Q = orc.quant_table(quality)
    host_px = strip.cpu().numpy()
    want_c, want_var, want_ties = orc.fwd_quant_plane(host_px, Q, adaptive, layout, nthreads=4)
    want_p, _ = orc.dequant_idct_plane(want_c, W, nrows, Q, adaptive, layout, want_var if adaptive else None, nthreads=4)
    plan.stats()
    var = torch.empty(nrows * W // 64, dtype=torch.float64, device=cx.dev) if adaptive else None
    got_c = plan.fwd_quant_dev(strip, layout, None, var)
    got_c = got_c[0] if isinstance(got_c, tuple) else got_c
    st = plan.stats()
    got_p = plan.dequant_idct_dev(got_c, W, nrows, layout, var)
    torch.cuda.synchronize()
    b0, b1 = r0 // 8 * (W // 8), (r0 + nrows) // 8 * (W // 8)
    ok = {"fwd_records": bool(np.array_equal(got_c.cpu().numpy(), want_c)),
          "inv_pixels": bool(np.array_equal(got_p.cpu().numpy(), want_p)),
          "tie_count": bool(st["near_ties"] == want_ties),
          "step_records": bool(np.array_equal(coef[b0:b1].cpu().numpy(), want_c)),
          "step_pixels": bool(np.array_equal(rec[r0:r0 + nrows].cpu().numpy(), want_p))}
    all_ok = all(ok.values())
    return cx.min_over_ranks(1.0 if all_ok else 0.0) == 1.0, ok


# ---------------------------------------------------------------------------------------------
# every BASELINE shape, device-resident and end to end (SURVEY.md 8d / north_star "each named shape")
# ---------------------------------------------------------------------------------------------
def shapes_block(cx, peak):
    """C1..C5 on the GPUs of this job.  Device-resident: CUDA events, median of 5, L2 flushed between repetitions
    for the shapes that fit in it, max over ranks.  End to end: the host-plane C-ABI calls on PINNED host buffers
    (H2D + kernels + D2H inside the timed region), host wall clock, max over ranks.
      C1  512x512, q50, one per GPU          C2  one 3840x2160 frame per GPU
      C3  one 7680x4320 4:2:0 frame per GPU (luma + 2 chroma planes, two tables, planes call)
      C4  4096 frames 1920x1080 sharded by frame (strong scaling)
      C5  one 65536x65536 image sharded by block-row range, quality sweep 10..95 (strong scaling)
    For C4 / C5 the end-to-end leg streams each rank's shard through the host calls in pieces of 1/8 of the job
    (one pinned window, reused), so host memory stays bounded; every byte of the shard crosses PCIe both ways."""
    import numpy as np

    from dct_b200 import api, sharding
    torch = cx.torch
    dev, world, rank = cx.dev, cx.world, cx.rank
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)

    def dev_ms(fn, reps=5, flush_l2=True):
        fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            if flush_l2:
                flush.fill_(1)
            cx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(cx.max_over_ranks(e0.elapsed_time(e1)))
        return float(np.median(ms))

    def entry(samples_total, ms, e2e_s=None, note=None):
        gpx = 2.0 * samples_total / ms / 1e6
        e = {"gpixel_s": gpx, "ms_fwd_plus_inv": ms, "frac_of_hbm_peak_per_gpu": 3.0 * gpx / world / peak}
        if e2e_s is not None:
            e["e2e_gpixel_s"] = 2.0 * samples_total / e2e_s / 1e9
            e["e2e_ms"] = e2e_s * 1e3
        if note:
            e["note"] = note
        return e

    d = api.dct_init(8)
    q50 = api.quant_init(8, 50, 0)
    plan50 = api.Plan(d, q50, cx.local)
    res = {"unit": "Gpixel/s, forward + inverse, whole job over all GPUs",
           "frac": "3 B per sample / time / n_gpus / measured copy peak",
           "e2e": "host-plane calls, pinned host buffers, H2D + kernels + D2H timed, host wall clock, max over ranks"}

    def one_plane_shape(Wp, Hp, plan, reps=5):
        px = torch.randint(0, 256, (Hp, Wp), dtype=torch.uint8, device=dev, generator=gen)
        coef = torch.empty(((Hp // 8) * (Wp // 8), 64), dtype=torch.int16, device=dev)
        out = torch.empty_like(px)

        def run():
            plan.fwd_quant_dev(px, coef_out=coef)
            plan.dequant_idct_dev(coef, Wp, Hp, pixels_out=out)

        ms = dev_ms(run, reps)
        h_px = torch.empty((Hp, Wp), dtype=torch.uint8).pin_memory()
        h_px.copy_(px.cpu())
        h_coef = torch.empty((coef.shape[0], 64), dtype=torch.int16).pin_memory()
        h_out = torch.empty((Hp, Wp), dtype=torch.uint8).pin_memory()

        def host():
            plan.fwd_quant_ptr(h_px.data_ptr(), Wp, Wp, Hp, h_coef.data_ptr(), 0)
            plan.dequant_idct_ptr(h_coef.data_ptr(), Wp, Hp, h_out.data_ptr(), Wp, 0)

        host()
        t = min(cx.wall(host) for _ in range(3))
        return ms, t

    # ---- C1 / C2: one small plane per GPU ------------------------------------------------------------
    ms, t = one_plane_shape(512, 512, plan50)
    res["C1 512x512 (one per GPU)"] = entry(512 * 512 * world, ms, t, "4 096 blocks: launch-bound")
    ms, t = one_plane_shape(3840, 2160, plan50)
    res["C2 3840x2160 frame (one per GPU)"] = entry(3840 * 2160 * world, ms, t, "single launches of 129 600 blocks")

    # ---- C3: one 8K 4:2:0 frame per GPU, luma q75 + Annex-K chroma table scaled by the reference's rule --------
    chroma = np.full((8, 8), 99.0)
    chroma[:4, :4] = [[17, 18, 24, 47], [18, 21, 26, 66], [24, 26, 56, 99], [47, 66, 99, 99]]
    q75, qc = api.quant_init(8, 75, 0), api.quant_init(8, 75, 0)
    api.set_quant_table(qc, np.clip(chroma * 0.5, 1.0, 255.0))
    pl, pc = api.Plan(d, q75, cx.local), api.Plan(d, qc, cx.local)
    shp = [(4320, 7680), (2160, 3840), (2160, 3840)]
    planes = [torch.randint(0, 256, s, dtype=torch.uint8, device=dev, generator=gen) for s in shp]
    recs = [torch.empty_like(p) for p in planes]
    coefs = [torch.empty((s[0] // 8 * (s[1] // 8), 64), dtype=torch.int16, device=dev) for s in shp]
    descs = (api.PlaneDesc * 3)()
    for i, (s, pp) in enumerate(zip(shp, (pl, pc, pc))):
        descs[i].plan = pp._h
        descs[i].pixels_in, descs[i].pixels_out = planes[i].data_ptr(), recs[i].data_ptr()
        descs[i].pitch, descs[i].width, descs[i].height = s[1], s[1], s[0]
        descs[i].coef, descs[i].variance = coefs[i].data_ptr(), None
    sp = api._stream_ptr(None)

    def c3():
        api._check(api._fwd_planes(descs, 3, api.NATURAL, sp))
        api._check(api._inv_planes(descs, 3, api.NATURAL, sp))

    ms = dev_ms(c3)
    hp = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in shp]
    hc = [torch.empty((c.shape[0], 64), dtype=torch.int16).pin_memory() for c in coefs]
    ho = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in shp]
    for a, b in zip(hp, planes):
        a.copy_(b.cpu())

    def c3_host():
        # the three planes' transfers overlap: asynchronous forms, one wait per plan
        for i, (s, pp) in enumerate(zip(shp, (pl, pc, pc))):
            if i == 2:
                pc.wait()                  # the two chroma planes share a plan (and its strip buffers)
            pp.fwd_quant_ptr_async(hp[i].data_ptr(), s[1], s[1], s[0], hc[i].data_ptr(), 0)
        pl.wait(), pc.wait()
        for i, (s, pp) in enumerate(zip(shp, (pl, pc, pc))):
            if i == 2:
                pc.wait()
            pp.dequant_idct_ptr_async(hc[i].data_ptr(), s[1], s[0], ho[i].data_ptr(), s[1], 0)
        pl.wait(), pc.wait()

    c3_host()
    t = min(cx.wall(c3_host) for _ in range(3))
    res["C3 7680x4320 4:2:0 frame (one per GPU)"] = entry(sum(s[0] * s[1] for s in shp) * world, ms, t,
                                                           "three planes, two tables, one planes call each way")
    del planes, recs, coefs, hp, hc, ho
    pl.close(), pc.close()

    # ---- C4: 4096 1080p frames, contiguous frame ranges per GPU ------------------------------------------
    Wf, Hf, F = 1920, 1080, 4096
    f0, f1 = sharding.frame_shard(F, rank, world)
    rows = (f1 - f0) * Hf
    px = torch.randint(0, 256, (rows, Wf), dtype=torch.uint8, device=dev, generator=gen)
    coef = torch.empty(((rows // 8) * (Wf // 8), 64), dtype=torch.int16, device=dev)
    out = torch.empty_like(px)

    def c4():
        plan50.fwd_quant_dev(px, coef_out=coef)
        plan50.dequant_idct_dev(coef, Wf, rows, pixels_out=out)

    ms = dev_ms(c4, reps=3, flush_l2=False)
    piece = (F // 8) * Hf                                   # pixel rows of 1/8 of the job
    n_pieces = rows // piece
    h_px = torch.empty((piece, Wf), dtype=torch.uint8).pin_memory()
    h_px.copy_(px[:piece].cpu())
    h_coef = [torch.empty((piece // 8 * (Wf // 8), 64), dtype=torch.int16).pin_memory() for _ in range(2)]
    h_out = torch.empty((piece, Wf), dtype=torch.uint8).pin_memory()
    plan_inv = api.Plan(d, q50, cx.local)

    def stream_pieces(pf, pi, Wp, prow, n, quality_layout=0):
        # forward of piece i+1 overlaps inverse of piece i: both PCIe directions busy
        pf.fwd_quant_ptr_async(h_px.data_ptr(), Wp, Wp, prow, h_coef[0].data_ptr(), quality_layout)
        for i in range(n):
            pf.wait()
            if i + 1 < n:
                pf.fwd_quant_ptr_async(h_px.data_ptr(), Wp, Wp, prow, h_coef[(i + 1) % 2].data_ptr(), quality_layout)
            pi.wait()
            pi.dequant_idct_ptr_async(h_coef[i % 2].data_ptr(), Wp, prow, h_out.data_ptr(), Wp, quality_layout)
        pi.wait()

    stream_pieces(plan50, plan_inv, Wf, piece, 1)
    t = cx.wall(lambda: stream_pieces(plan50, plan_inv, Wf, piece, n_pieces))
    res[f"C4 {F} frames 1920x1080, sharded by frame"] = entry(
        Wf * Hf * F, ms, t, f"{f1 - f0} frames per GPU; device-resident: one launch per direction; e2e: {n_pieces} pieces of {F // 8} frames "
                            f"per GPU through one pinned window, forward and inverse plans overlapped")
    del px, coef, out

    # ---- C5: one 65536^2 image, block-row ranges per GPU, quality sweep ------------------------------------
    S = 65536
    r0, r1 = sharding.block_row_shard(S, rank, world)
    px = torch.randint(0, 256, (r1 - r0, S), dtype=torch.uint8, device=dev, generator=gen)
    coef = torch.empty((((r1 - r0) // 8) * (S // 8), 64), dtype=torch.int16, device=dev)
    out = torch.empty_like(px)
    piece = S // 8                                          # pixel rows of 1/8 of the image
    n_pieces = (r1 - r0) // piece
    # the C4 window holds 1/8 of C4 = 1.06 Gpx; 1/8 of C5 is 0.54 Gpx: reuse the same pinned memory
    h_px = h_px.view(-1)[:piece * S].view(piece, S)
    h_coef = [c.view(-1)[:piece * S].view(-1, 64) for c in h_coef]
    h_out = h_out.view(-1)[:piece * S].view(piece, S)
    sweep = {}
    e2e_q = (10, 50, 75, 95)
    for quality in range(10, 100, 5):
        qq = api.quant_init(8, quality, 0)
        pq, pq_inv = api.Plan(d, qq, cx.local), api.Plan(d, qq, cx.local)

        def c5():
            pq.fwd_quant_dev(px, coef_out=coef)
            pq.dequant_idct_dev(coef, S, r1 - r0, pixels_out=out)

        sampler = ClockSampler(cx.local)          # this sweep runs minutes into the job: record what the clocks did
        sampler.start()
        t_c0 = time.perf_counter()
        ms = dev_ms(c5, reps=3, flush_l2=False)
        clk = sampler.stop(t_c0, time.perf_counter())
        st = pq.stats()
        t = None
        if quality in e2e_q:
            stream_pieces(pq, pq_inv, S, piece, 1)
            t = cx.wall(lambda: stream_pieces(pq, pq_inv, S, piece, n_pieces))
        e = entry(S * S, ms, t)
        e["replayed_fraction"] = cx.sum_over_ranks(st["replayed_blocks"]) / max(cx.sum_over_ranks(st["blocks"]), 1.0)
        e["exact_ties"] = int(cx.sum_over_ranks(st["near_ties"]))
        e["clocks_rank0"] = {k: clk.get(k) for k in ("sm_mhz", "power_w_max", "reasons")}
        sweep[f"q{quality}"] = e
        pq.close(), pq_inv.close()
        api.quant_free(qq)
    fr = [v["frac_of_hbm_peak_per_gpu"] for v in sweep.values()]
    res[f"C5 {S}x{S} image, block-row ranges, quality sweep"] = {
        "min_frac_over_sweep": min(fr), "max_frac_over_sweep": max(fr), "qualities": sweep,
        "note": f"{(r1 - r0)} pixel rows per GPU; replayed_fraction / exact_ties are forward + inverse passes together; "
                f"e2e at q{e2e_q}: {n_pieces} pieces of {piece} rows per GPU through one pinned window"}
    plan50.close(), plan_inv.close()
    api.dct_free(d), api.quant_free(q50), api.quant_free(q75), api.quant_free(qc)
    return res


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from dct_b200 import api

    assert torch.cuda.is_available() and api.device_count() > 0, "bench.py needs a CUDA device (no CPU fallback)"
    cx = Ctx()
    dev, local_rank = cx.dev, cx.local
    from dct_b200 import numa
    numa_info = numa.bind_to_gpu_node(local_rank) if world > 1 else {}   # pinned buffers on the GPU's socket
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    barrier = cx.barrier

    frames = args.frames
    rows = frames * H
    npx = rows * W
    nblocks = npx // 64
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    px = torch.randint(0, 256, (rows, W), dtype=torch.uint8, device=dev, generator=gen)
    coef = torch.empty((nblocks, 64), dtype=torch.int16, device=dev)
    rec = torch.empty_like(px)
    var = torch.empty(nblocks, dtype=torch.float64, device=dev) if args.adaptive else None

    dctx, qctx = api.dct_init(8), api.quant_init(8, args.quality, args.adaptive)
    plan = api.Plan(dctx, qctx, local_rank)

    def step():
        plan.fwd_quant_dev(px, args.layout, coef, var)
        plan.dequant_idct_dev(coef, W, rows, args.layout, var, rec)

    for _ in range(max(args.warmup, 3)):
        step()
    plan.stats()

    # ---- timed region: exactly K steps, device time, max over ranks -----------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.15)
    plan.profile(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    launches0 = plan.kernel_launches()
    barrier()
    t_wall0 = time.perf_counter()
    ev[0].record()
    for i in range(args.steps):
        plan.fwd_quant_dev(px, args.layout, coef, var)
        ev[2 * i + 1].record()
        plan.dequant_idct_dev(coef, W, rows, args.layout, var, rec)
        ev[2 * i + 2].record()
    barrier()
    t_wall1 = time.perf_counter()
    gpu_launches = plan.kernel_launches() - launches0        # counted by the library: K1, K3, K2, K3 per step
    plan.profile(False)
    total_ms = ev[0].elapsed_time(ev[2 * args.steps])
    fwd_ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps))
    inv_ms = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps))
    prof = plan.profile_fetch()
    stats = plan.stats()
    clocks = sampler.stop(t_wall0, t_wall1)

    total_ms_max = cx.max_over_ranks(total_ms)
    ms_per_step = total_ms_max / args.steps
    value = 2.0 * npx * world / (ms_per_step * 1e-3) / 1e9

    # ---- parity: a strip of this rank's output against the CPU oracle, MIN over ranks ----------------------
    parity_ok, parity_detail = parity_check(cx, plan, px, coef, rec, args.layout, args.quality, args.adaptive, frames)

    # ---- single-frame latency (BASELINE configs[1] as written: ONE 3840x2160 frame each way) --------------
    single = {}
    try:
        lat = []
        for i in range(min(frames, 24)):
            f_px, f_rec = px[i * H:(i + 1) * H], rec[i * H:(i + 1) * H]
            f_coef = coef[i * (H // 8) * (W // 8):(i + 1) * (H // 8) * (W // 8)]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            plan.fwd_quant_dev(f_px, args.layout, f_coef, None if var is None else var[:f_coef.shape[0]])
            plan.dequant_idct_dev(f_coef, W, H, args.layout, None if var is None else var[:f_coef.shape[0]], f_rec)
            e1.record()
            torch.cuda.synchronize()
            lat.append(e0.elapsed_time(e1) * 1e3)
        single = {"c2_fwd_plus_inv_us": statistics.median(lat[2:]), "min_us": min(lat[2:]),
                  "how": "one frame of the batch per measurement (cold in L2), CUDA events around its launches, median"}
    except Exception as e:  # pragma: no cover
        single = {"error": repr(e)}

    # ---- end to end: pinned host buffers, H2D + kernels + D2H inside the timed region -----------
    # Every step moves its batch host -> device -> host in both directions: forward reads pinned
    # pixels and leaves int16 records in pinned memory (where the untouched host entropy coder would
    # take them), inverse reads those records from pinned memory and leaves pixels in pinned memory.
    #   sync      : the two synchronous C-ABI calls back to back (forward is D2H-bound, inverse H2D-bound)
    #   pipelined : the asynchronous forms on two plans; forward of step i+1 overlaps inverse of step i,
    #               so both PCIe directions are busy (fill + drain are timed).
    e2e_frames = min(frames, 32)
    e_rows = e2e_frames * H
    e_px, e_nb = e_rows * W, e_rows * W // 64
    h_px = torch.empty((e_rows, W), dtype=torch.uint8).pin_memory()
    h_px.copy_(px[:e_rows].cpu())
    h_coef = [torch.empty((e_nb, 64), dtype=torch.int16).pin_memory() for _ in range(3)]
    h_rec = torch.empty((e_rows, W), dtype=torch.uint8).pin_memory()

    e2e = None
    if not args.adaptive:
        plan_inv = api.Plan(dctx, qctx, local_rank)

        def e2e_sync_step():
            plan.fwd_quant_ptr(h_px.data_ptr(), W, W, e_rows, h_coef[0].data_ptr(), args.layout)
            plan_inv.dequant_idct_ptr(h_coef[0].data_ptr(), W, e_rows, h_rec.data_ptr(), W, args.layout)

        def e2e_pipelined(n):
            # three record buffers decouple the two plans: forward i+1 is queued as soon as forward i is home,
            # inverse i as soon as inverse i-1 is; neither waits for the other beyond the data dependency
            plan.fwd_quant_ptr_async(h_px.data_ptr(), W, W, e_rows, h_coef[0].data_ptr(), args.layout)
            for i in range(n):
                plan.wait()                                            # records of step i are in h_coef[i % 3]
                if i + 1 < n:
                    plan.fwd_quant_ptr_async(h_px.data_ptr(), W, W, e_rows, h_coef[(i + 1) % 3].data_ptr(), args.layout)
                plan_inv.wait()                                        # inverse i-1 is home (frees h_rec)
                plan_inv.dequant_idct_ptr_async(h_coef[i % 3].data_ptr(), W, e_rows, h_rec.data_ptr(), W, args.layout)
            plan_inv.wait()

        e2e_sync_step()
        e2e_pipelined(1)
        n_e = max(args.e2e_steps, 1)
        t_sync = cx.wall(lambda: [e2e_sync_step() for _ in range(n_e)]) / n_e
        h_rec.zero_()
        t_pipe = cx.wall(lambda: e2e_pipelined(n_e)) / n_e
        want = plan.dequant_idct_dev(plan.fwd_quant_dev(px[:e_rows], args.layout), W, e_rows, args.layout)
        torch.cuda.synchronize()
        result_ok = bool(torch.equal(h_rec, want.cpu()))      # the step's result, read on the host
        best = min(t_sync, t_pipe)
        h2d_b, d2h_b = e_px + 2 * e_px, 2 * e_px + e_px
        ceiling = copy_ceiling(cx, h2d_b, d2h_b)
        e2e = {"value": 2.0 * e_px * world / best / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
               "ms_per_step": best * 1e3, "frames_per_step_per_gpu": e2e_frames, "steps": n_e,
               "mode": "pipelined" if t_pipe <= t_sync else "sync",
               "sync_value": 2.0 * e_px * world / t_sync / 1e9, "pipelined_value": 2.0 * e_px * world / t_pipe / 1e9,
               "gbs_per_direction": h2d_b * world / best / 1e9,
               "copy_ceiling_gbs": ceiling,
               "frac_of_ceiling": (h2d_b * world / best / 1e9) / ceiling["duplex_h2d"] if ceiling.get("duplex_h2d") else None,
               "pcie_link": sampler.pcie(), "host_cpus": os.cpu_count(),
               "timer": "host wall clock, barrier + synchronize on both sides, max over ranks",
               "numa_binding_rank0": numa_info,
               "result_matches_device_path": result_ok}
        # the same pipelined step with int8 records on the host side (lossless at this table, see
        # dct_cuda_plan_records_fit_i8): 2/3 of the PCIe bytes.  Reported next to the int16 figure, which stays
        # the headline (SURVEY 8d counts the records as int16).
        if plan.records_fit_i8:
            h_c8 = [torch.empty((e_nb, 64), dtype=torch.int8).pin_memory() for _ in range(3)]

            def e2e_pipelined_i8(n):
                plan.fwd_quant_i8_ptr_async(h_px.data_ptr(), W, W, e_rows, h_c8[0].data_ptr(), args.layout)
                for i in range(n):
                    plan.wait()
                    if i + 1 < n:
                        plan.fwd_quant_i8_ptr_async(h_px.data_ptr(), W, W, e_rows, h_c8[(i + 1) % 3].data_ptr(), args.layout)
                    plan_inv.wait()
                    plan_inv.dequant_idct_i8_ptr_async(h_c8[i % 3].data_ptr(), W, e_rows, h_rec.data_ptr(), W, args.layout)
                plan_inv.wait()

            e2e_pipelined_i8(1)
            h_rec.zero_()
            t_i8 = cx.wall(lambda: e2e_pipelined_i8(n_e)) / n_e
            e2e["int8_records"] = {"value": 2.0 * e_px * world / t_i8 / 1e9, "unit": UNIT, "ms_per_step": t_i8 * 1e3,
                                   "h2d_bytes_per_step": 2 * e_px, "d2h_bytes_per_step": 2 * e_px,
                                   "result_matches_device_path": bool(torch.equal(h_rec, want.cpu())),
                                   "note": "same step, records cross PCIe as int8 (every |q| <= 127 at this table)"}
            del h_c8
        plan_inv.close()
    del h_px, h_coef, h_rec

    # ---- roofline of the dominant kernel (K1), of K2, and of the whole step ---------------------------
    peak, peak_src = measured_peak()
    steps = max(args.steps, 1)
    k1_ms = prof["fwd_ms"] / steps        # K1 time per step (events around the K1 launch inside the library)
    k2_ms = prof["inv_ms"] / steps
    k3f_ms = max(fwd_ms / steps - k1_ms, 0.0)   # the rest of the forward phase: K3 and the launch gap
    k3i_ms = max(inv_ms / steps - k2_ms, 0.0)
    alg_bytes = 192.0 * nblocks
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
    step_achieved = 2.0 * alg_bytes / (total_ms / steps * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_fwd_quant_u8_tma (K1)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src + ", burst figure",
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k1_ms,
                "traffic": None, "traffic_source": None,
                "k2": {"kernel": "k_dequant_idct_u8_tma (K2)", "achieved": alg_bytes / (k2_ms * 1e-3) / 1e9,
                       "frac": alg_bytes / (k2_ms * 1e-3) / 1e9 / peak, "avg_launch_ms": k2_ms},
                "step_achieved": step_achieved, "step_frac": step_achieved / peak,
                "step_note": "2 x 192 B per block over the whole step (K1 + K3 + K2 + K3 and their gaps), this rank",
                "frac_of_8TBps_nominal": achieved / 8000.0}
    traffic_file = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            roofline["traffic"] = tj.get("dram_bytes_per_launch_scaled_to", {}).get(str(frames))
            roofline["traffic_source"] = tj.get("source", "profiles/k1_traffic.json") + " (an earlier ncu --set full capture of the same command, not this run)"
        except Exception:
            pass
    kernels = [{"name": "K1 k_fwd_quant_u8_tma", "us_per_step": k1_ms * 1e3},
               {"name": "K3f k_replay_fwd_lane (+ launch gap)", "us_per_step": k3f_ms * 1e3},
               {"name": "K2 k_dequant_idct_u8_tma", "us_per_step": k2_ms * 1e3},
               {"name": "K3i k_replay_inv_lane (+ launch gap)", "us_per_step": k3i_ms * 1e3}]

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 butterfly, f64 replay of near-ties (bit-exact int16/u8 results)", "data": "synthetic",
        "config": workload_config(args),
        "batch": {"frames_per_step_per_gpu": frames, "bytes_per_direction_per_gpu": 3 * npx},
        "fwd_gpixel_per_s_per_gpu": npx / (fwd_ms / args.steps * 1e-3) / 1e9,
        "inv_gpixel_per_s_per_gpu": npx / (inv_ms / args.steps * 1e-3) / 1e9,
        "e2e": e2e, "gpu_launches": gpu_launches,
        "launches_per_step": "K1 k_fwd_quant_u8_tma, K3 k_replay_fwd_lane, K2 k_dequant_idct_u8_tma, K3 k_replay_inv_lane",
        "kernels": kernels,
        "roofline": roofline, "clocks": clocks,
        "parity_ok": parity_ok, "parity_detail_rank0": parity_detail,
        "single_frame_us": single,
        "replay": {"blocks": stats["blocks"], "replayed_blocks": stats["replayed_blocks"],
                   "exact_ties": stats["near_ties"]},
    }
    del px, coef, rec
    plan.close()
    torch.cuda.empty_cache()

    if not args.no_shapes and not args.adaptive:
        try:
            line["shapes"] = shapes_block(cx, peak)
        except Exception as e:  # pragma: no cover  -- the contract line must still be printed
            line["shapes"] = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cf = args.cpu_frames or max(4, min(48, threads * 2))
        v, dt, kind = cpu_run(cf, threads, args.quality)
        cb = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "seconds": dt,
              "sample": f"{cf} of the same {W}x{H} frames, fwd+inv, {threads} pthreads over block rows, "
                        f"reference C sources at -O2 -std=c99"}
        v1, dt1, _ = cpu_run(1, 1, args.quality)
        cb["one_thread_O2"] = {"value": v1, "unit": UNIT, "cores": 1, "seconds": dt1, "sample": f"1 frame {W}x{H}, fwd+inv, -O2 -std=c99"}
        vs, dts, ks = cpu_run(0, 1, args.quality, which="ref_shipped", rows=720)
        if vs is not None:
            cb["one_thread_shipped_flags"] = {"value": vs, "unit": UNIT, "cores": 1, "seconds": dts,
                                              "sample": f"720 rows of a {W}-wide frame, fwd+inv, the reference's Justfile flags (-std=c99 -g, no -O)"}
        line["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(line), flush=True)
    api.dct_free(dctx), api.quant_free(qctx)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
