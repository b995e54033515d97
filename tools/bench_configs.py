"""Throughput of every BASELINE.json shape (configs[1..4]) on the GPUs of one box, device-resident.

north_star: "Throughput on synthetic images of each named shape is reported at 1/2/4/8 GPUs, in absolute numbers and
as a fraction of the HBM roofline."  bench.py's contract line covers configs[1] in batches; this tool walks all four
GPU shapes.  Exploratory (NOT the bench.py contract): same timing rules -- inputs larger than L2 or L2 flushed between
repetitions, CUDA events, barrier + max over ranks under torchrun -- no e2e leg.

  C2  one 3840x2160 frame                      1 GPU (a frame does not shard; N GPUs = N frames, one each)
  C3  one 7680x4320 4:2:0 frame                luma + two 3840x2160 chroma planes, two tables, planes call
  C4  4096 frames of 1920x1080                 contiguous frame ranges per GPU (strong scaling: total fixed)
  C5  one 65536x65536 image, q = 10..95        block-row ranges per GPU (strong scaling), K1 + K2 per quality

python tools/bench_configs.py                                   # 1 GPU
python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py --json out.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch
import torch.distributed as dist

from dct_b200 import api

CHROMA_K2 = np.full((8, 8), 99.0)
CHROMA_K2[:4, :4] = [[17, 18, 24, 47], [18, 21, 26, 66], [24, 26, 56, 99], [47, 66, 99, 99]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--c5-side", type=int, default=65536)
    ap.add_argument("--c4-frames", type=int, default=4096)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak = 6546.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def timed(fn, reps=a.reps, flush_l2=True):
        """median over reps of (barrier, [flush], events around fn); max over ranks"""
        fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            if flush_l2:
                flush.fill_(1)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms.append(float(t.item()))
        return float(np.median(ms))

    def entry(samples_total, ms, note=""):
        gpx = 2.0 * samples_total / ms / 1e6                       # forward + inverse
        return {"ms_fwd_plus_inv": ms, "gpixel_s": gpx, "frac_of_hbm_peak_per_gpu": 3.0 * gpx / world / peak, "note": note}

    d = api.dct_init(8)
    q50 = api.quant_init(8, 50, 0)
    plan50 = api.Plan(d, q50, local)
    res = {"n_gpus": world, "hbm_peak_gbs": peak, "unit": "Gpixel/s, forward + inverse, whole job",
           "frac": "3 B per sample / time / n_gpus / measured copy peak"}

    # ---- C2: one 4K frame per GPU -----------------------------------------------------------
    W, H = 3840, 2160
    px = torch.randint(0, 256, (H, W), dtype=torch.uint8, device=dev, generator=gen)
    coef = torch.empty(((H // 8) * (W // 8), 64), dtype=torch.int16, device=dev)
    out = torch.empty_like(px)

    def c2():
        plan50.fwd_quant_dev(px, coef_out=coef)
        plan50.dequant_idct_dev(coef, W, H, pixels_out=out)

    res["C2 3840x2160 frame (one per GPU)"] = entry(W * H * world, timed(c2), "single launches of 129 600 blocks: launch- and tail-bound, see C4 for the batched rate")

    # ---- C3: one 8K 4:2:0 frame per GPU, luma q75 + Annex-K chroma scaled by the reference's rule --
    q75 = api.quant_init(8, 75, 0)
    qc = api.quant_init(8, 75, 0)
    api.set_quant_table(qc, np.clip(CHROMA_K2 * 0.5, 1.0, 255.0))
    pl, pc = api.Plan(d, q75, local), api.Plan(d, qc, local)
    shapes = [(4320, 7680), (2160, 3840), (2160, 3840)]
    planes = [torch.randint(0, 256, s, dtype=torch.uint8, device=dev, generator=gen) for s in shapes]
    recs = [torch.empty_like(p) for p in planes]
    coefs = [torch.empty((s[0] // 8 * (s[1] // 8), 64), dtype=torch.int16, device=dev) for s in shapes]
    descs = (api.PlaneDesc * 3)()
    for i, (s, pp) in enumerate(zip(shapes, (pl, pc, pc))):
        descs[i].plan = pp._h
        descs[i].pixels_in, descs[i].pixels_out = planes[i].data_ptr(), recs[i].data_ptr()
        descs[i].pitch, descs[i].width, descs[i].height = s[1], s[1], s[0]
        descs[i].coef, descs[i].variance = coefs[i].data_ptr(), None
    sp = api._stream_ptr(None)

    def c3():
        api._check(api._fwd_planes(descs, 3, api.NATURAL, sp))
        api._check(api._inv_planes(descs, 3, api.NATURAL, sp))

    res["C3 7680x4320 4:2:0 frame (one per GPU)"] = entry(sum(s[0] * s[1] for s in shapes) * world, timed(c3), "three planes, two tables, one planes call each way")
    del planes, recs, coefs

    # ---- C4: 4096 1080p frames, contiguous frame ranges per GPU ------------------------------------
    W, H, F = 1920, 1080, a.c4_frames
    f0, f1 = F * rank // world, F * (rank + 1) // world
    rows = (f1 - f0) * H
    px = torch.randint(0, 256, (rows, W), dtype=torch.uint8, device=dev, generator=gen)
    coef = torch.empty(((rows // 8) * (W // 8), 64), dtype=torch.int16, device=dev)
    out = torch.empty_like(px)

    def c4():
        plan50.fwd_quant_dev(px, coef_out=coef)
        plan50.dequant_idct_dev(coef, W, rows, pixels_out=out)

    res[f"C4 {F} frames 1920x1080, sharded by frame"] = entry(W * H * F, timed(c4, flush_l2=False), f"{f1 - f0} frames per GPU, one launch per direction; the batch is far larger than L2")
    del px, coef, out

    # ---- C5: one 65536^2 image, block-row ranges per GPU, quality sweep ------------------------------
    S = a.c5_side
    r0, r1 = (S // 8) * rank // world * 8, (S // 8) * (rank + 1) // world * 8
    px = torch.randint(0, 256, (r1 - r0, S), dtype=torch.uint8, device=dev, generator=gen)
    coef = torch.empty((((r1 - r0) // 8) * (S // 8), 64), dtype=torch.int16, device=dev)
    out = torch.empty_like(px)
    sweep = {}
    for quality in range(10, 100, 5):
        qq = api.quant_init(8, quality, 0)
        pq = api.Plan(d, qq, local)

        def c5():
            pq.fwd_quant_dev(px, coef_out=coef)
            pq.dequant_idct_dev(coef, S, r1 - r0, pixels_out=out)

        ms = timed(c5, reps=3, flush_l2=False)
        st = pq.stats()
        t = torch.tensor([st["replayed_blocks"], st["blocks"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        e = entry(S * S, ms)
        e["replayed_fraction"] = float(t[0].item() / max(t[1].item(), 1.0))   # forward and inverse passes together
        sweep[f"q{quality}"] = e
        pq.close()
        api.quant_free(qq)
    res[f"C5 {S}x{S} image, block-row ranges, quality sweep"] = sweep

    if rank == 0:
        print(json.dumps(res))
        if a.json:
            with open(a.json, "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
