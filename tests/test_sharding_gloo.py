"""N>1 host logic on CPU: two gloo ranks shard a plane by block-row range and a batch by frame,
each processes only its shard, and the reassembled records / pixels equal the unsharded result.
(The arithmetic stand-in here is the CPU oracle -- there is no GPU on this box; on a B200 the same
ranges are fed to libdct_cuda, see tests/test_gpu_parity.py::test_multi_gpu_entry_point... and
bench.py --gpus N.)"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from dct_b200 import sharding
    from oracle import binding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = binding.load("oracle")
    H, W = 120, 256                                   # 15 block rows: does not divide evenly by 2
    px = orc.fill_xorshift(H, W, seed=99)
    Q = orc.quant_table(50)
    lo, hi = sharding.block_row_shard(H, rank, world)
    coef, _, ties = orc.fwd_quant_plane(px[lo:hi], Q, 0, 1)
    r0, r1 = sharding.record_range(W, lo, hi)
    assert coef.shape[0] == r1 - r0
    full = torch.zeros(((H // 8) * (W // 8), 64), dtype=torch.int32)
    full[r0:r1] = torch.from_numpy(coef.astype(np.int32))
    dist.all_reduce(full)                             # disjoint ranges: sum == concatenation (test plumbing only)
    t = torch.tensor([ties], dtype=torch.int64)
    dist.all_reduce(t)
    # frames: 5 frames of 16 rows over 2 ranks
    flo, fhi = sharding.frame_shard(5, rank, world)
    frames = torch.zeros(5, dtype=torch.int64)
    frames[flo:fhi] = 1
    dist.all_reduce(frames)
    if rank == 0:
        want, _, want_ties = orc.fwd_quant_plane(px, Q, 0, 1)
        np.save(os.path.join(out_dir, "ok.npy"),
                np.array([np.array_equal(full.numpy(), want), int(t.item()) == want_ties, bool((frames == 1).all())]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_block_row_and_frame_sharding(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    ok = np.load(tmp_path / "ok.npy")
    assert ok.all(), ok


def test_ranges_cover_exactly_once():
    from dct_b200 import sharding
    for total in (0, 1, 7, 8, 4096, 8191):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.contiguous_range(total, r, world)
                assert 0 <= lo <= hi <= total and hi - lo in (total // world, total // world + 1)
                seen += list(range(lo, hi))
            assert seen == list(range(total))
    assert sharding.block_row_shard(65536, 3, 8) == (3 * 8192, 4 * 8192)      # config 5: 1024 block rows per GPU
    assert sharding.record_range(65536, 8192, 16384) == (1024 * 8192, 2048 * 8192)
    with pytest.raises(ValueError):
        sharding.block_row_shard(12, 0, 2)
