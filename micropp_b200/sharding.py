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


def balanced_ranges(costs, nproc: int) -> list[tuple[int, int]]:
    """Cost-aware contiguous shards: [begin, end) per rank such that the largest per-rank cost is minimal.

    `costs[g]` is what Gauss point g cost in the last macro step -- ``micropp<3>::get_cost`` (CG iterations,
    src/homogenize.cpp:127-131), what the reference's test/mpi-load-balance.cpp:56-73 shows to differ by 3-5x between
    linear and non-linear Gauss points.  Ranges stay contiguous (a macro code scatters strains by slice; the FE state
    of a Gauss point lives on its GPU, so a re-shard moves `u_n`/`vars_n` through write_restart/read_restart).  With
    equal costs the result is `gp_range` (remainders to the low ranks).  Exact: binary search on the bottleneck value
    with a greedy feasibility sweep, O(ngp log(sum))."""
    w = [max(float(c), 0.0) + 1.0 for c in costs]   # +1: a linear GP still costs its set-up, and empty ranks are avoided
    ngp = len(w)
    nproc = max(1, int(nproc))
    if ngp == 0:
        return [(0, 0)] * nproc

    def cuts(limit):
        out, acc, begin = [], 0.0, 0
        for g, c in enumerate(w):
            if acc + c > limit and g > begin:
                out.append((begin, g))
                begin, acc = g, 0.0
            acc += c
        out.append((begin, ngp))
        return out

    lo, hi = max(w), sum(w)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if len(cuts(mid)) <= nproc:
            hi = mid
        else:
            lo = mid
    parts = cuts(hi)
    # fewer parts than ranks: split the longest ranges so that every rank owns work when ngp >= nproc
    while len(parts) < nproc:
        k = max(range(len(parts)), key=lambda q: parts[q][1] - parts[q][0])
        b, e = parts[k]
        if e - b < 2:
            break
        parts[k:k + 1] = [(b, (b + e) // 2), ((b + e) // 2, e)]
    parts += [(ngp, ngp)] * (nproc - len(parts))
    return parts
