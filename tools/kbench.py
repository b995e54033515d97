#!/usr/bin/env python
"""Kernel-level timing of one step (K1, K3 forward, K2, K3 inverse) on the bench batch -- a quick loop for kernel work.

    python tools/kbench.py [--quality 50] [--adaptive 0] [--layout 0] [--frames 64] [--steps 20] [--W 3840 --H 2160]

Prints one JSON line: microseconds per launch (CUDA events inside the library around K1 / K2; the K3 figures are
the rest of each phase, launch gap included), fractions of the measured copy peak, replay statistics, and a hash of
the outputs so that two builds / switches (e.g. DCT_CUDA_NO_TMA=1) can be compared for bit-equality.
Exploratory tool, not the bench.py contract.
"""
import argparse
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from dct_b200 import api


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quality", type=int, default=50)
    ap.add_argument("--adaptive", type=int, default=0)
    ap.add_argument("--layout", type=int, default=0)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--W", type=int, default=3840)
    ap.add_argument("--H", type=int, default=2160)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    W, rows = a.W, a.H * a.frames
    gen = torch.Generator(device=dev).manual_seed(1234)
    px = torch.randint(0, 256, (rows, W), dtype=torch.uint8, device=dev, generator=gen)
    nblocks = rows * W // 64
    coef = torch.empty((nblocks, 64), dtype=torch.int16, device=dev)
    rec = torch.empty_like(px)
    var = torch.empty(nblocks, dtype=torch.float64, device=dev) if a.adaptive else None
    d, q = api.dct_init(8), api.quant_init(8, a.quality, a.adaptive)
    plan = api.Plan(d, q, 0)
    for _ in range(3):
        plan.fwd_quant_dev(px, a.layout, coef, var)
        plan.dequant_idct_dev(coef, W, rows, a.layout, var, rec)
    plan.stats()
    plan.profile(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * a.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(a.steps):
        plan.fwd_quant_dev(px, a.layout, coef, var)
        ev[2 * i + 1].record()
        plan.dequant_idct_dev(coef, W, rows, a.layout, var, rec)
        ev[2 * i + 2].record()
    torch.cuda.synchronize()
    plan.profile(False)
    fwd = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(a.steps)) / a.steps * 1e3
    inv = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(a.steps)) / a.steps * 1e3
    pf = plan.profile_fetch()
    st = plan.stats()
    k1, k2 = pf["fwd_ms"] / a.steps * 1e3, pf["inv_ms"] / a.steps * 1e3
    peak = 6546.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs = lambda us: 192.0 * nblocks / us / 1e3
    h = hashlib.sha1()
    h.update(coef.cpu().numpy().tobytes())
    h.update(rec.cpu().numpy().tobytes())
    print(json.dumps({"tag": a.tag, "quality": a.quality, "adaptive": a.adaptive, "layout": a.layout, "shape": [W, a.H, a.frames],
                      "k1_us": round(k1, 1), "k3f_us": round(fwd - k1, 1), "k2_us": round(k2, 1), "k3i_us": round(inv - k2, 1),
                      "step_us": round(fwd + inv, 1), "gpixel_s": round(2.0 * rows * W / (fwd + inv) / 1e3, 1),
                      "k1_frac": round(gbs(k1) / peak, 4), "k2_frac": round(gbs(k2) / peak, 4),
                      "step_frac": round(2 * gbs(fwd + inv) / peak, 4),
                      "replayed_frac": round(st["replayed_blocks"] / max(st["blocks"], 1), 5), "ties": st["near_ties"],
                      "sha1": h.hexdigest()[:16]}), flush=True)
    plan.close()


if __name__ == "__main__":
    main()
