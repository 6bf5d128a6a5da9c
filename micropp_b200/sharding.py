"""Distribution of independent macro Gauss points over ranks / GPUs (no data-path collective).

The reference shards Gauss points in its MPI test drivers, never inside the library:
``ngp_per_mpi = ngp / nproc + (ngp % nproc > rank)`` (test/multi-gpu-mpi.cpp:60, test/benchmark-sc2019.cpp:69),
one ``micropp<3>`` object per rank and ``gpu_id = mpi_rank % ngpus`` (src/micropp.cpp:77-81).  These helpers
state the same rule as contiguous ranges so that a macro code can scatter strains / gather stresses by slice.
"""
from __future__ import annotations


def gp_count(ngp: int, nproc: int, rank: int) -> int:
    """Number of Gauss points owned by `rank` (test/multi-gpu-mpi.cpp:60)."""
    return ngp // nproc + (1 if ngp % nproc > rank else 0)


def gp_range(ngp: int, nproc: int, rank: int) -> tuple[int, int]:
    """Contiguous [begin, end) of global Gauss-point ids owned by `rank`; remainders go to the low ranks."""
    begin = sum(gp_count(ngp, nproc, r) for r in range(rank))
    return begin, begin + gp_count(ngp, nproc, rank)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Device/wall time of a multi-rank step = the slowest rank (bench.py contract)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
