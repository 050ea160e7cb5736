"""Shard arithmetic for the multi-GPU paths (pure host logic, no CUDA).

The path has no exchange step: every 4x4 block is independent (GoofyTC/goofy_tc.h:1514-1524
just walks tiles), so multi-GPU is a static partition -- by texture for batches, by horizontal
strips of whole block rows for one large image -- and nothing is reduced or gathered.
`torch.distributed` is used by bench.py only for the barrier and the max-over-ranks timing.
"""
from __future__ import annotations

from dataclasses import dataclass


def strip_partition(height: int, n_shards: int, shard: int) -> tuple[int, int]:
    """Python twin of goofy_b200_strip_partition (capi.cu): block rows [first, first+count)."""
    rows = height // 4
    if n_shards <= 0 or not (0 <= shard < n_shards):
        return 0, 0
    first = rows * shard // n_shards
    nxt = rows * (shard + 1) // n_shards
    return first, nxt - first


def batch_partition(n_images: int, n_shards: int, shard: int) -> range:
    """Contiguous range of texture indices for `shard` (sizes differ by at most one)."""
    if n_shards <= 0 or not (0 <= shard < n_shards):
        return range(0)
    return range(n_images * shard // n_shards, n_images * (shard + 1) // n_shards)


@dataclass(frozen=True)
class Strip:
    shard: int
    first_row: int      # pixel row
    rows: int           # pixel rows
    src_offset: int     # bytes into the source image
    dst_offset: int     # bytes into the block output
    dst_bytes: int


def strips(width: int, height: int, stride: int, n_shards: int) -> list[Strip]:
    out = []
    for g in range(n_shards):
        first, count = strip_partition(height, n_shards, g)
        out.append(Strip(g, first * 4, count * 4, first * 4 * stride, first * (width // 4) * 8,
                         count * (width // 4) * 8))
    return out
