"""CPU tests of the N>1 host logic: Gauss-point sharding (world_size-2 gloo) and the max-over-ranks reduction.

GPs are independent RVEs, so the multi-GPU path has no data-path collective: each rank owns a
contiguous range and runs its own solver; only the timing reduction and the final gather use the process group.
Here every rank evaluates its range with the CPU oracle and the gathered result must equal the serial one.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from micropp_b200.sharding import gp_count, gp_range, max_over_ranks, sum_over_ranks


def test_ranges_partition_like_the_reference_drivers():
    for ngp in (1, 7, 8, 10, 4096, 4099):
        for nproc in (1, 2, 3, 4, 8):
            counts = [gp_count(ngp, nproc, r) for r in range(nproc)]
            assert sum(counts) == ngp
            assert counts == [ngp // nproc + (1 if ngp % nproc > r else 0) for r in range(nproc)]
            ends = [gp_range(ngp, nproc, r) for r in range(nproc)]
            assert ends[0][0] == 0 and ends[-1][1] == ngp
            assert all(ends[r][1] == ends[r + 1][0] for r in range(nproc - 1))


def _worker(rank, world, port, ngp, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import orcpy as O
    el, dm = (0, 3e7, 0.3, 0, 0, 0), (2, 1e7, 0.3, 0, 0, 1e5)
    b, e = gp_range(ngp, world, rank)
    eps = np.random.default_rng(1234).uniform(-2e-3, 2e-3, (ngp, 6))  # the same global list on every rank
    m = O.OrcMicropp(dict(size=(4, 4, 4), type=1, geo_params=(0.3, 0, 0, 0), materials=[dm, el, el], ngp=e - b,
                          lin_stress=False))
    for g in range(b, e):
        m.set_strain(g - b, eps[g])
    m.homogenize()
    mine = torch.zeros((ngp, 6), dtype=torch.float64)
    for g in range(b, e):
        mine[g] = torch.from_numpy(m.get_stress(g - b))
    dist.all_reduce(mine)  # disjoint ranges: the sum is the gather
    t = max_over_ranks(1.0 + rank, dist)
    n = sum_over_ranks(float(e - b), dist)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), mine.numpy())
        np.save(os.path.join(out_dir, "scalars.npy"), np.array([t, n]))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_serial(tmp_path):
    from oracle import orcpy as O
    ngp, world = 5, 2
    port = 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ngp, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    t, n = np.load(tmp_path / "scalars.npy")
    assert t == 2.0 and n == ngp
    el, dm = (0, 3e7, 0.3, 0, 0, 0), (2, 1e7, 0.3, 0, 0, 1e5)
    eps = np.random.default_rng(1234).uniform(-2e-3, 2e-3, (ngp, 6))
    m = O.OrcMicropp(dict(size=(4, 4, 4), type=1, geo_params=(0.3, 0, 0, 0), materials=[dm, el, el], ngp=ngp,
                          lin_stress=False))
    for g in range(ngp):
        m.set_strain(g, eps[g])
    m.homogenize()
    want = np.array([m.get_stress(g) for g in range(ngp)])
    assert np.array_equal(got, want)  # GP independence: bit-identical however the batch is split


def _halo_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from micropp_b200.slab import exchange_planes, plane_range
    nz, plane = 7, 5
    z0, z1 = plane_range(nz, world, rank)
    glob = torch.arange(nz * plane, dtype=torch.float64).view(nz, plane)   # the "p" every rank should agree on
    lo, hi = int(rank > 0), int(rank + 1 < world)
    local = torch.full((z1 - z0 + lo + hi, plane), -1.0, dtype=torch.float64)
    local[lo:lo + z1 - z0] = glob[z0:z1]                                      # owned planes only
    rlo, rhi = torch.empty(plane, dtype=torch.float64), torch.empty(plane, dtype=torch.float64)
    exchange_planes(dist, rank, world, local[lo].clone(), local[lo + z1 - z0 - 1].clone(), rlo, rhi)
    if lo:
        local[0] = rlo
    if hi:
        local[-1] = rhi
    ok = torch.equal(local, glob[z0 - lo:z1 + hi])
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(out_dir, "halo_ok.npy"), t.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_pattern_gloo(tmp_path, world):
    port = 31000 + (os.getpid() % 2000) + world
    mp.spawn(_halo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / "halo_ok.npy")[0] == 1.0


def test_balanced_ranges_cost_aware():
    """(f) rank 4: contiguous cost-aware shards (test/mpi-load-balance.cpp:56-73: 25 % non-linear GPs cost 3-5x)."""
    from micropp_b200.sharding import balanced_ranges, gp_range
    # equal costs => the reference drivers' rule
    for ngp, nproc in ((16, 4), (10, 3), (7, 8), (10, 4), (9, 4), (1, 3), (4096, 8)):
        got = balanced_ranges([5] * ngp, nproc)
        assert got == [gp_range(ngp, nproc, r) for r in range(nproc)], (ngp, nproc, got)
    # unequal costs: contiguous cover, bottleneck optimal (checked against brute force), parts even
    import itertools
    rng = np.random.default_rng(3)
    for _ in range(30):
        ngp, nproc = int(rng.integers(3, 11)), int(rng.integers(2, 5))
        costs = [int(c) for c in rng.integers(0, 50, ngp)]
        got = balanced_ranges(costs, nproc)
        assert len(got) == nproc and got[0][0] == 0 and got[-1][1] == ngp
        assert all(got[i][1] == got[i + 1][0] for i in range(nproc - 1))
        w = [c + 1.0 for c in costs]
        best = min(max(sum(w[b:e]) for b, e in zip((0,) + cut, cut + (ngp,)))
                   for cut in itertools.combinations_with_replacement(range(ngp + 1), nproc - 1))
        assert max(sum(w[b:e]) for b, e in got) <= best * (1 + 1e-9), (costs, nproc, got)
    # the reference's imbalance scenario: the first quarter of the GPs is non-linear and 4x as expensive
    costs = [400] * 16 + [100] * 48
    parts = balanced_ranges(costs, 4)
    assert len(parts) == 4 and parts[0][0] == 0 and parts[-1][1] == 64
    assert all(parts[i][1] == parts[i + 1][0] for i in range(3))
    load = [sum(costs[b:e]) for b, e in parts]
    naive = [sum(costs[gp_range(64, 4, r)[0]:gp_range(64, 4, r)[1]]) for r in range(4)]
    assert max(load) < 0.55 * max(naive)          # 6400 on rank 0 with the plain rule, about 2800 balanced
    assert max(load) <= 1.15 * sum(costs) / 4 + max(costs)
