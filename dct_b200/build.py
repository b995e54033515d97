"""Builds dct_b200/libdct_cuda.so in-tree with nvcc for sm_100a (no torch involved).

    python -m dct_b200.build            # incremental
    python -m dct_b200.build --force

The host context code is compiled as ISO C99 by gcc (like the reference), the kernels and the
C-ABI shim by nvcc; everything links into one shared library with the CUDA runtime linked
statically, so a plain C program only needs `-ldct_cuda`.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdct_cuda.so")
BUILD = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_SOURCES = ["fwd_quant.cu", "dequant_idct.cu", "replay_f64.cu", "rle.cu", "generic_n.cu", "planar.cu", "narrow.cu", "tma_host.cu", "shim.cu", "shim_frames.cu", "shim_peer.cu", "shim_blocks.cu"]
C_SOURCES = ["host_context.c"]
HEADERS = ["butterfly.cuh", "fast_core.cuh", "kernels.cuh", "plan.cuh", "band_tables.h", "tma.cuh", "replay_lane.cuh"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    inc = os.path.join(ROOT, "include")
    pub_headers = [os.path.join(inc, h) for h in ("dct.h", "quantization.h", "utils.h", "dct_cuda.h")]
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + pub_headers
    objs, cmds = [], []
    for src in CU_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(BUILD, src + ".o")
        if force or _newer(o, [s] + hdrs):
            cmd = [NVCC, "-std=c++17", "-O3", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-ffp-contract=off",
                   "-I" + inc, "-I" + CSRC, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            cmds.append(cmd)
        objs.append(o)
    if verbose or len(cmds) < 2:          # -v: keep ptxas' output in source order
        for cmd in cmds:
            subprocess.check_call(cmd)
    else:                                 # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as pool:
            list(pool.map(subprocess.check_call, cmds))
    for src in C_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(BUILD, src + ".o")
        if force or _newer(o, [s] + pub_headers):
            subprocess.check_call(["gcc", "-std=c99", "-ffp-contract=off", "-O2", "-Wall", "-Wextra", "-Werror",
                                   "-pedantic", "-fPIC", "-I" + inc, "-c", s, "-o", o])
        objs.append(o)
    if force or _newer(OUT, objs):
        subprocess.check_call([NVCC, "-shared", *ARCH, "-cudart", "static", "-o", OUT, *objs, "-lm", "-lpthread"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
