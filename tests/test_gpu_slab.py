"""-m gpu: one RVE split into z-slabs (BASELINE configs[4]).  On a single GPU the slabs live in one process and the
halo exchange / all-reduce are device copies, which exercises the decomposition itself (ownership of planes and
element layers, halo consistency, slab-local sums + scalar tails); the NCCL transport of the same code path is
covered by tools/slab_bench.py under `gpurun --gpus N` and by tests/test_sharding.py (gloo) for the message pattern.

Bar: the slab solution equals the single-domain solution of the same kernels to rounding (the only difference is the
summation order of the dot products), iteration counts within +-1, stress within 1e-8 of the reference CPU path."""
import numpy as np
import pytest

from common import CASES, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("dims,nslabs", [((10, 9, 12), 2), ((10, 9, 12), 3), ((12, 12, 12), 4), ((8, 8, 8), 1)])
def test_slab_equals_single_domain(mpp, refpy, dims, nslabs, exchange):
    """exchange="peer": halo planes pulled from the neighbour's vector and dot products summed through mailboxes with
    device-side epoch flags (the NVLink P2P path; here all slabs sit on one GPU, so the "peer" pointers are local);
    "nccl": the send/recv + all-reduce transport (device copies in this single-process mode)."""
    from micropp_b200.slab import SlabRVE
    kw = dict(size=dims, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"])
    eps = np.array([1.0e-3, -0.4e-3, 0.2e-3, 0.6e-3, -0.3e-3, 0.1e-3])
    one = mpp.Micropp3(mpp.default_params(**kw))
    one.set_strain(0, eps)
    one.homogenize()
    s1, c1 = one.get_stress(0), one.get_cost(0)
    u1 = one.get_u(0, 1).reshape(-1, 3)

    rve = SlabRVE(kw, nslabs=nslabs, exchange=exchange)
    out = rve.homogenize(eps)
    assert rve.peer_error() == 0
    assert out["converged"] and abs(out["cg_its"] - c1) <= 1
    assert relerr(out["stress"], s1) < 1e-10
    assert relerr(rve.get_u(), u1) < 1e-8
    assert rve.exchanges > 0 and rve.allreduces > 0
    rve.close()

    r = refpy.RefMicropp(refpy.default_params(**kw))
    r.set_strain(0, eps)
    r.homogenize()
    assert relerr(out["stress"], r.get_stress(0)) < 1e-8
    assert abs(out["cg_its"] - r.get_cost(0)) <= 1


def test_slab_damage_first_step(mpp):
    # history-free first load step of a damage RVE: tangents and residuals across the cuts
    from micropp_b200.slab import SlabRVE
    kw = dict(size=(9, 9, 11), lin_stress=False, calc_ctan_lin=False, nr_max_its=10, **CASES["damage_sphere"])
    eps = np.array([0.012, 0, 0, 0, 0, 0.0])
    one = mpp.Micropp3(mpp.default_params(**kw))
    one.set_strain(0, eps)
    one.homogenize()
    rve = SlabRVE(kw, nslabs=3)
    out = rve.homogenize(eps)
    assert rve.peer_error() == 0
    assert out["converged"] == one.has_converged(0)
    assert abs(out["cg_its"] - one.get_cost(0)) <= out["newton_its"]
    assert relerr(out["stress"], one.get_stress(0)) < 1e-8
    rve.close()


def test_plane_ranges():
    from micropp_b200.slab import plane_range
    for nz in (8, 50, 200):
        for r in (1, 2, 3, 4, 8):
            rr = [plane_range(nz, r, s) for s in range(r)]
            assert rr[0][0] == 0 and rr[-1][1] == nz and all(a[1] == b[0] for a, b in zip(rr[:-1], rr[1:]))


def test_slab_host_is_cxx_and_refuses_history(mpp):
    """The product path of the slab mode is the C++ host behind micropp3x_slab_* (include/micropp_b200_ext.h); it keeps
    no internal variables between calls, so a damage RVE may be solved once (virgin state) and a second call is refused
    loudly instead of silently using 'no history'."""
    from micropp_b200.slab import SlabRVE, SlabRVECxx
    kw = dict(size=(8, 8, 9), lin_stress=False, calc_ctan_lin=False, nr_max_its=10, **CASES["damage_sphere"])
    rve = SlabRVE(kw, nslabs=2)
    assert isinstance(rve, SlabRVECxx) and rve.op == 0        # a damage phase: every slab on its own ELL matrix
    out = rve.homogenize(np.array([0.012, 0, 0, 0, 0, 0.0]))
    one = mpp.Micropp3(mpp.default_params(**kw))
    one.set_strain(0, np.array([0.012, 0, 0, 0, 0, 0.0]))
    one.homogenize()
    assert out["converged"] == one.has_converged(0) and relerr(out["stress"], one.get_stress(0)) < 1e-8
    assert rve.launch_count() > 0
    with pytest.raises(RuntimeError):
        rve.homogenize(np.array([0.013, 0, 0, 0, 0, 0.0]))
    rve.close()
    kw = dict(size=(8, 8, 9), lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"])
    rve = SlabRVE(kw, nslabs=3)
    assert rve.op == 3                                        # all-elastic: the implicit operator on every slab
    a = rve.homogenize(np.array([1e-3, 0, 0, 0, 0, 0.0]))
    b = rve.homogenize(np.array([1e-3, 0, 0, 0, 0, 0.0]))    # elastic: repeatable, bit for bit
    assert np.array_equal(a["stress"], b["stress"]) and a["cg_its"] == b["cg_its"]
    rve.close()


@pytest.mark.parametrize("nslabs", [0, 1, 2, 3])
def test_cg_residual_history_vs_oracle(mpp, nslabs):
    """First-K residual history of the DPCG solve: |z| at the head of every iteration (the quantity src/ell.cpp:93-94
    tests) of the single-domain product (nslabs = 0) and of the z-slab mode against the oracle's restatement of
    ell_solve_cgpd on the matrix and right-hand side the oracle assembles itself.  Differences are rounding (summation
    order of 243-term rows and of the three dot products per iteration) amplified along the recurrence."""
    from micropp_b200.slab import SlabRVE
    from oracle import orcpy as O
    dims, K = (10, 9, 12), 64
    kw = dict(size=dims, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"])
    eps = np.array([1.0e-3, -0.4e-3, 0.2e-3, 0.6e-3, -0.3e-3, 0.1e-3])
    p = O.OrcProblem(kw)
    u = p.set_displ_bc(eps)
    b, _ = p.assembly_rhs(u)
    _, its, want = O.ell_solve_cgpd_hist(*dims, p.assembly_mat(u), b, K)
    if nslabs == 0:
        m = mpp.Micropp3(mpp.default_params(**kw))
        m.cg_history(K)
        m.set_strain(0, eps)
        m.homogenize()
        got, cost = m.cg_history_read(0, K), m.get_cost(0)
        m.close()
    else:
        rve = SlabRVE(kw, nslabs=nslabs)
        rve.cg_history(K)
        out = rve.homogenize(eps)
        assert rve.peer_error() == 0
        got, cost = rve.cg_history_read(K), out["cg_its"]
        for s in range(1, nslabs):        # every slab saw the same globally summed values
            assert np.array_equal(rve.cg_history_read(K, slab=s), got)
        rve.close()
    assert cost == its and len(got) == len(want) == min(K, its + 1)
    d = np.abs(got - want) / want
    print(f"residual history, nslabs={nslabs}: {len(got)} values, relative difference to the oracle: first 16 "
          f"{d[:16].max():.1e}, iterations 16-31 {d[16:32].max():.1e}, all {d.max():.1e}")
    # rounding only, but CG amplifies it once the Krylov basis loses orthogonality: a numpy restatement of the SAME
    # algorithm that sums with pairwise instead of sequential additions differs from the oracle by 1e-14 over the first
    # 20 iterations, 1e-11 at 30 and 2e-6 at the 40th (last) one -- the product shows the same profile
    assert d[:16].max() < 1e-12 and d[:32].max() < 1e-8 and d.max() < 2e-5
