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
    of a Gauss point lives on its GPU, so a re-shard moves `u_n`/`vars_n` through write_restart/read_restart).

    Equal costs reproduce `gp_range` exactly (the reference drivers' rule, remainders to the low ranks).  Otherwise:
    (1) the minimal bottleneck B by binary search with a greedy feasibility sweep, (2) a left-to-right pass that ends
    rank r at the position closest to the cumulative target (r+1)/nproc of the total, subject to every part <= B and
    to the rest still fitting the remaining ranks -- so the parts are even, not just bounded."""
    w = [max(float(c), 0.0) + 1.0 for c in costs]   # +1: a linear GP still costs its set-up, and empty ranks are avoided
    ngp = len(w)
    nproc = max(1, int(nproc))
    if ngp == 0:
        return [(0, 0)] * nproc
    if max(w) == min(w):
        return [gp_range(ngp, nproc, r) for r in range(nproc)]

    def nparts(limit, begin):
        """greedy count of parts of cost <= limit that cover [begin, ngp)"""
        n, acc = 1, 0.0
        for g in range(begin, ngp):
            if acc + w[g] > limit and acc > 0.0:
                n, acc = n + 1, 0.0
            acc += w[g]
        return n if begin < ngp else 0

    lo, hi = max(w), sum(w)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if nparts(mid, 0) <= nproc:
            hi = mid
        else:
            lo = mid
    B = hi * (1.0 + 1e-12)
    total = sum(w)
    pre = [0.0]
    for c in w:
        pre.append(pre[-1] + c)
    parts, begin = [], 0
    for r in range(nproc):
        left = nproc - r - 1                      # ranks after this one
        if r == nproc - 1 or begin >= ngp:
            parts.append((begin, ngp if r == nproc - 1 else begin))
            begin = parts[-1][1]
            continue
        target = total * (r + 1) / nproc
        best, best_d = None, None
        for end in range(begin + 1, ngp - left + 1 if ngp - begin > left else ngp + 1):
            if pre[end] - pre[begin] > B:
                break
            if nparts(B, end) > left:
                continue
            d = abs(pre[end] - target)
            if best is None or d < best_d:
                best, best_d = end, d
        if best is None:
            best = min(begin + 1, ngp)
        parts.append((begin, best))
        begin = best
    return parts
