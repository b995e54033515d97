#!/usr/bin/env python
"""bench.py -- Gpixel/s of fused forward DCT+quantise and dequantise+IDCT on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 3840x2160 8-bit grayscale frames, quality 50, adaptive off,
NATURAL layout.  One STEP = one pass of the hot path over one batch of `--frames` synthetic frames:
K1 (u8 pixels -> int16 records) followed by K2 (records -> u8 pixels), both through libdct_cuda's
C ABI.  The batch (64 frames = 531 MB of pixels + 1.06 GB of records) is larger than the 126 MB
L2, so no flush is needed between iterations.  (Larger batches gain ~1 % and run into the software
power cap within a 100-step region: 128 frames 1 692, 256 frames 1 698 Gpixel/s against 1 677.)  `value` counts every pixel once per direction:
    value = 2 * frames * W * H * n_gpus / step_time           [Gpixel/s, device-resident]
`e2e` is the same metric through the host-plane calls with PINNED HOST buffers, the H2D and D2H
copies inside the timed region.  `roofline` is for K1 (the dominant kernel): 192 algorithmic bytes
per 8x8 block (64 B of pixels in, 128 B of records out) over the K1 phase's own
CUDA-event time, against the measured copy bandwidth in MEASURED_PEAKS.json.  `cpu_baseline` times
the reference's own C code (oracle/_ref, else the oracle port) on this box's host cores.

Multi-GPU: every rank owns its own batch (weak scaling), no data-path collective; torch.distributed
(NCCL) is used only for the barrier and the max-over-ranks of the device time.
`--impl reference` times the reference's CPU implementation on all host threads instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
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
    ap.add_argument("--quality", type=int, default=QUALITY, help="non-default values are exploratory, not the headline config")
    return ap.parse_args()


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
def cpu_checker():
    from oracle import binding
    if binding.have_ref():
        return binding.load("ref"), "reference"
    return binding.load("oracle"), "port"


def cpu_run(frames, threads, repeats=1):
    """fwd+inv of `frames` 4K frames through the reference's block functions; returns (Gpx/s, seconds)."""
    import numpy as np
    lib, kind = cpu_checker()
    rng = np.random.default_rng(1)
    px = rng.integers(0, 256, size=(frames * H, W), dtype=np.uint8)
    Q = lib.quant_table(QUALITY)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        coef, _, _ = lib.fwd_quant_plane(px, Q, 0, 0, nthreads=threads)
        lib.dequant_idct_plane(coef, W, frames * H, Q, 0, 0, None, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return 2.0 * frames * W * H / best / 1e9, best, kind


def reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = args.cpu_frames or max(2, min(16, threads // 2))
    cpu_run(1, threads)                      # warm the page cache / thread pool once
    times = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        v, dt, kind = cpu_run(frames, threads)
        if i >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    value = 2.0 * frames * W * H / t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{W}x{H} 8-bit grayscale frames, q{QUALITY}, adaptive off: forward DCT+quantise "
                               f"then dequantise+IDCT (BASELINE configs[1])",
                   "frames_per_step": frames, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{frames} frames of {W}x{H} per step, fwd+inv, {threads} pthreads over block rows, -O2 -std=c99"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from dct_b200 import api

    assert torch.cuda.is_available() and api.device_count() > 0, "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from dct_b200 import numa
    numa_info = numa.bind_to_gpu_node(local_rank) if world > 1 else {}   # pinned buffers on the GPU's socket
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = 2.0 * npx * world / (ms_per_step * 1e-3) / 1e9

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

        def timed(fn):
            barrier()
            t0 = time.perf_counter()
            fn()
            barrier()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        e2e_sync_step()
        e2e_pipelined(1)
        n_e = max(args.e2e_steps, 1)
        t_sync = timed(lambda: [e2e_sync_step() for _ in range(n_e)]) / n_e
        h_rec.zero_()
        t_pipe = timed(lambda: e2e_pipelined(n_e)) / n_e
        want = plan.dequant_idct_dev(plan.fwd_quant_dev(px[:e_rows], args.layout), W, e_rows, args.layout)
        torch.cuda.synchronize()
        result_ok = bool(torch.equal(h_rec, want.cpu()))      # the step's result, read on the host
        best = min(t_sync, t_pipe)
        e2e = {"value": 2.0 * e_px * world / best / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": e_px + 2 * e_px, "d2h_bytes_per_step": 2 * e_px + e_px,
               "ms_per_step": best * 1e3, "frames_per_step_per_gpu": e2e_frames, "steps": n_e,
               "mode": "pipelined" if t_pipe <= t_sync else "sync",
               "sync_value": 2.0 * e_px * world / t_sync / 1e9, "pipelined_value": 2.0 * e_px * world / t_pipe / 1e9,
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
            t_i8 = timed(lambda: e2e_pipelined_i8(n_e)) / n_e
            e2e["int8_records"] = {"value": 2.0 * e_px * world / t_i8 / 1e9, "unit": UNIT, "ms_per_step": t_i8 * 1e3,
                                   "h2d_bytes_per_step": 2 * e_px, "d2h_bytes_per_step": 2 * e_px,
                                   "result_matches_device_path": bool(torch.equal(h_rec, want.cpu())),
                                   "note": "same step, records cross PCIe as int8 (every |q| <= 127 at this table)"}
        plan_inv.close()

    # ---- roofline of the dominant kernel (K1) -----------------------------------------------------
    peak, peak_src = measured_peak()
    k1_ms = prof["fwd_ms"] / max(args.steps, 1)        # K1 time per step
    k2_ms = prof["inv_ms"] / max(args.steps, 1)
    alg_bytes = 192.0 * nblocks
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_fwd_quant_u8 (K1)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src + ", burst figure",
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k1_ms,
                "traffic": None,
                "k2": {"kernel": "k_dequant_idct_u8 (K2)", "achieved": alg_bytes / (k2_ms * 1e-3) / 1e9,
                       "frac": alg_bytes / (k2_ms * 1e-3) / 1e9 / peak, "avg_launch_ms": k2_ms},
                "frac_of_8TBps_nominal": achieved / 8000.0}
    traffic_file = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch_scaled_to", {}).get(str(frames))
        except Exception:
            pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 butterfly, f64 replay of near-ties (bit-exact int16/u8 results)", "data": "synthetic",
        "config": {"workload": f"{W}x{H} 8-bit grayscale frames, q{args.quality}, adaptive {'on' if args.adaptive else 'off'}: "
                               f"forward DCT+quantise then dequantise+IDCT (BASELINE configs[1])",
                   "frames_per_step_per_gpu": frames, "layout": "zigzag" if args.layout else "natural",
                   "l2": "inputs larger than L2 (531 MB pixels + 1062 MB records per direction), no flush",
                   "sharding": "one batch per GPU, no collective on the data path"},
        "fwd_gpixel_per_s_per_gpu": npx / (fwd_ms / args.steps * 1e-3) / 1e9,
        "inv_gpixel_per_s_per_gpu": npx / (inv_ms / args.steps * 1e-3) / 1e9,
        "e2e": e2e, "gpu_launches": gpu_launches,
        "launches_per_step": "K1 k_fwd_quant_u8, K3 k_replay_fwd, K2 k_dequant_idct_u8, K3 k_replay_inv",
        "roofline": roofline, "clocks": clocks,
        "replay": {"blocks": stats["blocks"], "replayed_blocks": stats["replayed_blocks"],
                   "exact_ties": stats["near_ties"]},
    }

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cf = args.cpu_frames or max(4, min(48, threads * 2))
        v, dt, kind = cpu_run(cf, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "seconds": dt,
                                "sample": f"{cf} of the same {W}x{H} frames, fwd+inv, {threads} pthreads over block rows, "
                                          f"reference C sources at -O2 -std=c99"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    plan.close()
    api.dct_free(dctx), api.quant_free(qctx)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
