"""Multi-GPU partitioning of the pair workloads (one process per GPU, no collective).

Independent scan pairs and all-pairs loop-closure candidates shard by contiguous ranges of
the pair index; every rank runs the same single-GPU kernel on its range.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of rank's share; sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(int(n_items), world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def triangle_pair_count(n_rows: int) -> int:
    return n_rows * (n_rows - 1) // 2


def triangle_pair(q: int, n_rows: int) -> Tuple[int, int]:
    """Host mirror of the device's linear-index -> (i, j), i < j, row-major."""
    i, rem = 0, q
    while rem >= n_rows - 1 - i:
        rem -= n_rows - 1 - i
        i += 1
    return i, i + 1 + rem
