"""Device-side throughput of the planar front / back end (csrc/planar.cu) on a batch of RGB frames.

Exploratory tool (NOT the bench.py contract): CUDA-event timing of the colour-conversion kernels and of the
whole RGB -> records -> RGB chain on resident data, against the measured HBM copy peak.
Usage: python tools/bench_planar.py [--width 7680 --height 4320 --frames 8 --reps 10] [--json out.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch

from dct_b200 import api


def timed(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.fill_(1)                      # > L2: the next launch starts cold
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=7680)
    ap.add_argument("--height", type=int, default=4320)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    W, H = a.width, a.height * a.frames       # frames stacked vertically: one tall image (heights are multiples of 16)
    peak = 6546.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    g = torch.Generator(device="cuda").manual_seed(1)
    rgb = api._alloc_rgb(H, W, "cuda")
    rgb.copy_(torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device="cuda", generator=g))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    geo = api.frame420_geometry(W, H)
    y = torch.empty((geo.y_height, geo.y_width), dtype=torch.uint8, device="cuda")
    cb = torch.empty((geo.c_height, geo.c_width), dtype=torch.uint8, device="cuda")
    cr = torch.empty_like(cb)
    out = api._alloc_rgb(H, W, "cuda")
    sp = api._stream_ptr(None)
    import ctypes as C

    def fwd():
        api._check(api._rgb_to_ycc(0, rgb.data_ptr(), rgb.stride(0), C.byref(geo), y.data_ptr(), y.stride(0), cb.data_ptr(),
                                   cr.data_ptr(), cb.stride(0), sp))

    def inv():
        api._check(api._ycc_to_rgb(0, y.data_ptr(), y.stride(0), cb.data_ptr(), cr.data_ptr(), cb.stride(0), C.byref(geo),
                                   out.data_ptr(), out.stride(0), sp))

    px = W * H
    res = {"width": a.width, "height": a.height, "frames": a.frames, "pixels": px, "hbm_peak_gbps": peak}
    for name, fn in (("rgb_to_ycbcr420", fwd), ("ycbcr420_to_rgb", inv)):
        ms = timed(fn, a.reps, flush)
        gbps = 4.5 * px / ms / 1e6
        res[name] = {"ms": ms, "gpixel_s": px / ms / 1e6, "gb_s": gbps, "frac_of_hbm_peak": gbps / peak}

    # whole chain on resident data: RGB -> planes -> records (K1 x3) and back (K2 x3 -> RGB)
    d, ql, qc = api.dct_init(8), api.quant_init(8, 75, 0), api.quant_init(8, 75, 0)
    chroma = np.full((8, 8), 99.0)
    chroma[:4, :4] = [[17, 18, 24, 47], [18, 21, 26, 66], [24, 26, 56, 99], [47, 66, 99, 99]]
    api.set_quant_table(qc, np.clip(chroma * 0.5, 1.0, 255.0))
    pl, pc = api.Plan(d, ql), api.Plan(d, qc)
    ky = torch.empty((y.numel() // 64, 64), dtype=torch.int16, device="cuda")
    kcb = torch.empty((cb.numel() // 64, 64), dtype=torch.int16, device="cuda")
    kcr = torch.empty_like(kcb)

    def encode():
        fwd()
        pl.fwd_quant_dev(y, api.ZIGZAG, coef_out=ky)
        pc.fwd_quant_dev(cb, api.ZIGZAG, coef_out=kcb)
        pc.fwd_quant_dev(cr, api.ZIGZAG, coef_out=kcr)

    def decode():
        pl.dequant_idct_dev(ky, geo.y_width, geo.y_height, api.ZIGZAG, pixels_out=y)
        pc.dequant_idct_dev(kcb, geo.c_width, geo.c_height, api.ZIGZAG, pixels_out=cb)
        pc.dequant_idct_dev(kcr, geo.c_width, geo.c_height, api.ZIGZAG, pixels_out=cr)
        inv()

    for name, fn in (("encode_rgb_to_records", encode), ("decode_records_to_rgb", decode)):
        ms = timed(fn, a.reps, flush)
        # bytes: conversion 4.5 B/px + transform of 1.5 samples/px at 3 B/sample
        res[name] = {"ms": ms, "gpixel_s": px / ms / 1e6, "gb_s": 9.0 * px / ms / 1e6,
                     "frac_of_hbm_peak": 9.0 * px / ms / 1e6 / peak}
    print(json.dumps(res))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
