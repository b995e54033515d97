#!/usr/bin/env python
"""What the host side of the box can move: bare pinned-memory cudaMemcpyAsync, N GPUs at once.

    python tools/pcie_ceiling.py                                          # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py

One process per GPU (as bench.py runs), all ranks copy at the same time between barriers; GB/s summed over ranks on
the slowest rank's time.  Byte counts are those of one bench.py e2e step (796 MB up, 796 MB down per GPU).
Variants: H2D only, D2H only, both directions at once; the H2D source in ordinary pinned memory or write-combined
pinned memory; copies in one piece or in 32-Mpx strips (96 MB) as the host pipeline of libdct_cuda issues them.
Also prints the box: CPUs, NUMA nodes, memory, PCIe link of every GPU, `nvidia-smi topo -m`.
This is the denominator of bench.py's e2e.frac_of_ceiling; exploratory tool, not the bench contract.
"""
import ctypes
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def red(x, op):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=op)
        return float(t.item())

    def wall(fn):
        barrier()
        t0 = time.perf_counter()
        fn()
        barrier()
        return red(time.perf_counter() - t0, dist.ReduceOp.MAX)

    nbytes = 3 * 32 * 3840 * 2160          # one direction of one e2e step
    strip = 3 * (32 << 20)                 # the host pipeline's strips: 32 Mpx x 3 B
    h_up = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_up.fill_(1)
    d_up = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_dn = torch.full((nbytes,), 2, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    # write-combined pinned memory for the H2D source (cudaHostAllocWriteCombined = 4)
    rt = None
    for name in ("libcudart.so.12", "libcudart.so"):     # the runtime torch already loaded
        try:
            rt = ctypes.CDLL(name)
            break
        except OSError:
            continue
    assert rt is not None, "no CUDA runtime library to call cudaMemcpyAsync through"
    wc_ptr = ctypes.c_void_p()
    wc = None
    try:
        if rt.cudaHostAlloc(ctypes.byref(wc_ptr), ctypes.c_size_t(nbytes), 4) == 0:
            ctypes.memset(wc_ptr, 1, nbytes)
            wc = wc_ptr
    except Exception:
        wc = None

    def memcpy_async(dst, src, n, kind, stream):
        # kind: 1 = H2D, 2 = D2H
        rt.cudaMemcpyAsync(ctypes.c_void_p(dst), ctypes.c_void_p(src), ctypes.c_size_t(n), kind, ctypes.c_void_p(stream.cuda_stream))

    reps = 4

    def up(pieces=False, src=None):
        base = src if src is not None else h_up.data_ptr()
        for _ in range(reps):
            if pieces:
                for o in range(0, nbytes, strip):
                    memcpy_async(d_up.data_ptr() + o, base + o, min(strip, nbytes - o), 1, s_up)
            else:
                memcpy_async(d_up.data_ptr(), base, nbytes, 1, s_up)

    def down(pieces=False):
        for _ in range(reps):
            if pieces:
                for o in range(0, nbytes, strip):
                    memcpy_async(h_dn.data_ptr() + o, d_dn.data_ptr() + o, min(strip, nbytes - o), 2, s_dn)
            else:
                memcpy_async(h_dn.data_ptr(), d_dn.data_ptr(), nbytes, 2, s_dn)

    res = {"n_gpus": world, "bytes_per_direction_per_gpu": nbytes, "unit": "GB/s per direction, summed over GPUs"}

    def measure(name, fn):
        fn()
        barrier()
        t = min(wall(fn) for _ in range(3)) / reps
        res[name] = round(nbytes * world / t / 1e9, 2)

    measure("h2d_only", up)
    measure("d2h_only", down)
    measure("duplex", lambda: (up(), down()))
    measure("duplex_in_96MB_strips", lambda: (up(True), down(True)))
    if wc is not None:
        measure("h2d_only_write_combined_source", lambda: up(False, wc.value))
        measure("duplex_write_combined_source", lambda: (up(False, wc.value), down()))
    res["per_gpu_duplex"] = round(res["duplex"] / world, 2)

    if rank == 0:
        box = {"cpus": os.cpu_count()}
        try:
            box["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        except OSError:
            box["numa_nodes"] = None
        try:
            box["mem_total_gb"] = round(int(open("/proc/meminfo").readline().split()[1]) / 1e6, 1)
        except Exception:
            pass
        for key, cmd in (("pcie", ["nvidia-smi", "--query-gpu=index,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv,noheader"]),
                         ("topo", ["nvidia-smi", "topo", "-m"])):
            try:
                box[key] = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout.strip().splitlines()
            except Exception as e:
                box[key] = repr(e)
        res["box"] = box
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
