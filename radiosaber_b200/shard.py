"""Multi-GPU plumbing: cells are independent (the reference runs one cell per process,
single-cell-with-interference.h:74), so a batch is block-partitioned over ranks with no collective
on the data path; the only exchange is one sum-reduce of the per-slice integer totals at the end of
a run (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_cells(total_cells: int, world: int, rank: int):
    """Contiguous block of cells for `rank`: (first_cell, n_cells). Blocks differ by at most one."""
    base, rem = divmod(int(total_cells), int(world))
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def reduce_stats(stats, dst: int = 0):
    """Sum the uint64 [4][S] per-slice totals over all ranks onto `dst` (returns the tensor; only
    meaningful on dst).  `stats` is a torch int64 tensor (bit pattern of the uint64 totals: two's
    complement addition wraps identically) on the device the process group works on."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(stats, dst=dst, op=dist.ReduceOp.SUM)
    return stats


def stats_from_state(cum_bytes: np.ndarray, cum_rbs: np.ndarray, ue_to_slice: np.ndarray, n_slices: int) -> np.ndarray:
    """Host statement of rs_get_stats(): uint64 [4][S] = sum bytes, sum RBs, sum q, sum q*q with
    q = bytes >> 10, over every cell and UE of a slice (used by the CPU tests)."""
    out = np.zeros((4, n_slices), dtype=np.uint64)
    q = cum_bytes >> np.uint64(10)
    for s in range(n_slices):
        m = ue_to_slice == s
        out[0, s] = cum_bytes[:, m].sum(dtype=np.uint64)
        out[1, s] = cum_rbs[:, m].sum(dtype=np.uint64)
        out[2, s] = q[:, m].sum(dtype=np.uint64)
        out[3, s] = (q[:, m] * q[:, m]).sum(dtype=np.uint64)
    return out
