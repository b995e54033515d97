"""Host-side partitioning of the path across GPUs (no data-path collective).

Blocks are independent (src/dct.c:52-77 and src/quantization.c:113-126 read only their own block
and immutable tables), so a batch shards by frame and a single large image by block-row range.
These are the same contiguous ranges dct_cuda_*_multi uses inside libdct_cuda (shim.cu:run_sharded).
"""
from __future__ import annotations


def contiguous_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of `total` units for `rank` of `world`: contiguous, balanced to within one unit."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    return total * rank // world, total * (rank + 1) // world


def frame_shard(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """BASELINE config 4: contiguous frame ranges per GPU."""
    return contiguous_range(n_frames, rank, world)


def block_row_shard(height: int, rank: int, world: int) -> tuple[int, int]:
    """BASELINE config 5: contiguous 8-pixel block-row ranges; returns PIXEL rows [lo, hi)."""
    if height % 8:
        raise ValueError("height must be a multiple of 8")
    lo, hi = contiguous_range(height // 8, rank, world)
    return lo * 8, hi * 8


def record_range(width: int, row_lo: int, row_hi: int) -> tuple[int, int]:
    """The block-major record indices covered by pixel rows [row_lo, row_hi) of a `width`-wide plane."""
    bw = width // 8
    return (row_lo // 8) * bw, (row_hi // 8) * bw
