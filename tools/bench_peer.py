"""NVLink peer path (dct_cuda_*_peer): throughput against the peers' share of the rows, on the GPUs visible.

Exploratory tool (NOT the bench.py contract).  The plane and the records live on GPU 0; the time is taken with CUDA
events on GPU 0's stream, which the call orders after every peer's shard.
Usage: python tools/bench_peer.py [--gpus N] [--frames 64] [--json out.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch

from dct_b200 import api


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--exact", action="store_true", help="a table below 1.0: every block takes the fp64 exact path")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    n = a.gpus or api.device_count()
    W, H = 3840, 2160 * a.frames
    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda:0").manual_seed(3)
    px = torch.randint(0, 256, (H, W), dtype=torch.uint8, device="cuda:0", generator=g)
    coef = torch.empty(((H // 8) * (W // 8), 64), dtype=torch.int16, device="cuda:0")
    out = torch.empty_like(px)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    d = api.dct_init(8)
    q = api.quant_init(8, 50, 0)
    if a.exact:
        t = api.matrix_of(q.contents.quant_matrix, 8).copy()
        t[7, 7] = 0.75
        api.set_quant_table(q, t)
    plans = [api.Plan(d, q, dev) for dev in range(n)]
    res = {"gpus": n, "pixels": W * H, "exact_path": bool(a.exact), "default_share_per_peer": {"fwd": float(api._peer_share(plans[0]._h, n, 1)), "inv": float(api._peer_share(plans[0]._h, n, 0))}, "sweep": []}
    totals = [0.0, 0.05, 0.10, 0.15, 0.20, 0.25, 0.30, 0.40, 0.50] if n > 1 else [0.0]
    if a.exact and n > 1:
        totals = [0.0, 0.25, (n - 1) / n, 0.75]
    for tot in totals:
        share = [0.0] + [tot / (n - 1)] * (n - 1) if n > 1 else [0.0]
        row = {"peers_total_share": tot}
        for name, fn in (("fwd", lambda: api.fwd_quant_peer(plans, px, coef_out=coef, share=share)),
                         ("inv", lambda: api.dequant_idct_peer(plans, coef, W, H, pixels_out=out, share=share))):
            fn()
            torch.cuda.synchronize()
            ms = []
            for _ in range(a.reps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            t = float(np.median(ms))
            row[name] = {"ms": t, "gpixel_s": W * H / t / 1e6}
        res["sweep"].append(row)
        print(json.dumps(row), flush=True)
    for p in plans:
        p.close()
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
