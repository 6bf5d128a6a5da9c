"""-m gpu: the implicit operator of all-elastic RVEs (no per-slot ELL matrix; DPCG applies the table of distinct
row blocks to 8 right-hand sides per thread) against (a) the explicit per-slot-matrix path of the same library,
which must agree BIT FOR BIT (same row values, same FMA order, same reduction tree), and (b) the reference."""
import os

import numpy as np
import pytest

from common import CASES, relerr

pytestmark = pytest.mark.gpu


def run(mod_cls, params, eps, implicit=None, kernel=None):
    """One homogenize() of a fresh object; implicit / kernel select the operator through the environment
    (MICROPP_IMPLICIT: 0 = one assembled ELL matrix per slot; MICROPP_IMP_KERNEL: 0 = k_spmv_dot_imp, the table-driven
    multi-RHS kernel, 3 = k_spmv_dot_tmac + k_spmv_fix (TMA load and TMA store) when nx is even -- both inside the
    three-kernel DPCG loop (MICROPP_RESIDENT=0); unset = the default: the cluster-resident DPCG kernel whenever the RVE
    fits a cluster, tests/test_gpu_resident.py).  The residual of an all-elastic RVE is b = -A u through the implicit
    operator by default; kernel 0 keeps the element loop (MICROPP_RHS_OPERATOR=0) for the bit-wise comparison."""
    env = {}
    if implicit is not None:
        env["MICROPP_IMPLICIT"] = "1" if implicit else "0"
    if kernel is not None:
        env["MICROPP_IMP_KERNEL"] = str(kernel)
        env["MICROPP_RESIDENT"] = "0"
    if kernel == 0:
        env["MICROPP_RHS_OPERATOR"] = "0"   # the element loop of assembly_rhs, as the explicit path (bit-wise comparison)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        g = mod_cls(params)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    ngp = eps.shape[0]
    for gp in range(ngp):
        g.set_strain(gp, eps[gp])
    g.homogenize()
    sig = np.array([g.get_stress(gp) for gp in range(ngp)])
    cost = [g.get_cost(gp) for gp in range(ngp)]
    conv = [g.has_converged(gp) for gp in range(ngp)]
    return g, sig, cost, conv


DIMS = [((12, 12, 12), 19), ((9, 11, 10), 8), ((14, 9, 8), 3), ((30, 7, 6), 2), ((19, 12, 5), 5), ((3, 3, 3), 1)]


@pytest.mark.parametrize("dims,ngp", DIMS)
def test_implicit_simple_kernel_equals_explicit_bitwise(mpp, dims, ngp):
    rng = np.random.default_rng(42)
    eps = rng.uniform(-1e-3, 1e-3, (ngp, 6))
    kw = dict(size=dims, ngp=ngp, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"])
    gi, si, ci, vi = run(mpp.Micropp3, mpp.default_params(**kw), eps, implicit=True, kernel=0)
    ge, se, ce, ve = run(mpp.Micropp3, mpp.default_params(**kw), eps, implicit=False)
    assert gi.implicit_rows() >= 1 and ge.implicit_rows() == 0
    assert ci == ce and vi == ve
    assert np.array_equal(si, se)
    for gp in (0, ngp - 1):
        assert np.array_equal(gi.get_u(gp), ge.get_u(gp))


@pytest.mark.parametrize("kernel", [None, 3])
@pytest.mark.parametrize("dims,ngp", DIMS + [((16, 16, 16), 9), ((20, 10, 12), 2), ((30, 30, 30), 3), ((40, 18, 70), 2)])
def test_implicit_tma_kernels_equal_explicit(mpp, dims, ngp, kernel):
    """kernel None: the default path (cluster-resident DPCG where the RVE fits); 3: the three-kernel loop with
    k_spmv_dot_tmac + k_spmv_fix when nx is even, else the table-driven kernel -- said on stderr.  Ap is bit-identical (test_operator_application), p.Ap is summed in another fixed order, so the DPCG path
    differs by rounding.  DPCG stops at |z| < 1e-5 |z0| and amplifies rounding differences by about 1/tolerance x condition:
    cubic meshes stay below 1e-9, strongly anisotropic ones (dx != dy != dz) reach 1e-7 -- the same size as the
    difference between the assembled-matrix path and the reference CPU path on those meshes."""
    rng = np.random.default_rng(43)
    eps = rng.uniform(-1e-3, 1e-3, (ngp, 6))
    kw = dict(size=dims, ngp=ngp, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"])
    gi, si, ci, vi = run(mpp.Micropp3, mpp.default_params(**kw), eps, implicit=True, kernel=kernel)
    ge, se, ce, ve = run(mpp.Micropp3, mpp.default_params(**kw), eps, implicit=False)
    assert vi == ve and all(abs(a - b) <= 1 for a, b in zip(ci, ce))
    tol = 1e-9 if len(set(dims)) == 1 else 1e-4
    for gp in range(ngp):
        assert relerr(si[gp], se[gp]) < tol
    # same strain on every slot => bit-identical results on every slot (deterministic reductions)
    eps1 = np.tile(eps[:1], (ngp, 1))
    _, s1, c1, _ = run(mpp.Micropp3, mpp.default_params(**kw), eps1, implicit=True, kernel=kernel)
    assert all(np.array_equal(s1[0], s1[gp]) for gp in range(ngp)) and len(set(c1)) == 1


def test_implicit_homogeneous_single_row(mpp):
    kw = dict(size=(6, 6, 6), ngp=1, type=0, materials=[(0, 3.0e7, 0.25, 0, 0, 0)] * 3, lin_stress=False,
              calc_ctan_lin=False)
    g, sig, cost, conv = run(mpp.Micropp3, mpp.default_params(**kw), np.array([[1e-3, 0, 0, 0, 0, 0]]), implicit=True)
    assert g.implicit_rows() == 3      # the three pure-material row blocks are always present
    # homogeneous strain: sigma = C eps exactly (test/benchmark-elastic.cpp material)
    lam, mu = 0.25 * 3.0e7 / (1.25 * 0.5), 3.0e7 / 2.5
    assert relerr(sig[0][:3], [(lam + 2 * mu) * 1e-3, lam * 1e-3, lam * 1e-3]) < 1e-10


@pytest.mark.parametrize("dims,ngp", [((10, 10, 10), 11), ((7, 8, 9), 4)])
def test_implicit_vs_reference(mpp, refpy, dims, ngp):
    rng = np.random.default_rng(7)
    eps = rng.uniform(-1e-3, 1e-3, (ngp, 6))
    kw = dict(size=dims, ngp=ngp, lin_stress=False, calc_ctan_lin=True, **CASES["elastic_sphere"])
    g, sg, cg, vg = run(mpp.Micropp3, mpp.default_params(**kw), eps, implicit=True)
    r, sr, cr, vr = run(refpy.RefMicropp, refpy.default_params(**kw), eps)
    assert vg == vr
    for gp in range(ngp):
        assert relerr(sg[gp], sr[gp]) < 1e-8          # north-star tolerance
        assert abs(cg[gp] - cr[gp]) <= 1              # CG iterations (one Newton step each)
    assert relerr(g.ctan_lin(), r.ctan_lin()) < 1e-8  # the 6 unit-strain solves of the constructor


def test_use_A0_on_elastic_rve_matches(mpp):
    # use_A0 on an all-elastic RVE: A0 == the implicit operator, results must not change
    rng = np.random.default_rng(3)
    eps = rng.uniform(-1e-3, 1e-3, (5, 6))
    kw = dict(size=(8, 8, 8), ngp=5, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"])
    _, s0, c0, _ = run(mpp.Micropp3, mpp.default_params(**kw), eps, implicit=True)
    _, s1, c1, _ = run(mpp.Micropp3, mpp.default_params(use_A0=True, its_with_A0=1, **kw), eps, implicit=True)
    assert c0 == c1 and np.array_equal(s0, s1)


@pytest.mark.parametrize("dims", [(14, 9, 8), (12, 12, 12), (30, 7, 6), (19, 12, 5), (3, 3, 3), (11, 10, 13),
                                  (40, 11, 7), (66, 5, 6), (4, 4, 4), (30, 30, 30), (16, 14, 21), (24, 12, 9),
                                  (10, 6, 14), (50, 8, 8)])
def test_operator_application_three_ways(mpp, refpy, dims):
    """A p through the assembled ELL matrix, the simple implicit kernel and the tiled implicit kernel: identical
    bits; and equal to the reference's ell_mvp on the reference's own assembled matrix."""
    kw = dict(size=dims, ngp=1, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"])
    g = mpp.Micropp3(mpp.default_params(**kw))
    rng = np.random.default_rng(5)
    p = rng.uniform(-1.0, 1.0, g.nndim).reshape(-1, 3)
    nx, ny, nz = dims
    idx = np.arange(nx * ny * nz)
    i, j, k = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    bnd = (i == 0) | (i == nx - 1) | (j == 0) | (j == ny - 1) | (k == 0) | (k == nz - 1)
    p[bnd] = 0.0   # the search direction vanishes on the boundary (identity rows, b = 0 there)
    p = p.reshape(-1)
    y0, d0 = g.apply_operator(p, op=0)
    y1, d1 = g.apply_operator(p, op=3, kernel=0)
    assert g.implicit_kernel() == (3 if nx % 2 == 0 else 0)   # k_spmv_dot_tmac whenever TMA can address the rows
    bad = np.nonzero(y0 != y1)[0]
    if bad.size:   # diagnostics: which of the two is unstable, and which agrees with the reference
        y0b, _ = g.apply_operator(p, op=0)
        y1b, _ = g.apply_operator(p, op=3, kernel=0)
        r = refpy.RefMicropp(refpy.default_params(**kw))
        yr = refpy.ell_mvp(*dims, r.assembly_mat(np.zeros(g.nndim)), p)
        inner = ~np.repeat(bnd, 3)
        raise AssertionError(dict(n_bad=int(bad.size), first=bad[:4].tolist(), y0_rep=int(np.sum(y0 != y0b)),
                                  y1_rep=int(np.sum(y1 != y1b)), y0b_vs_y1b=int(np.sum(y0b != y1b)),
                                  err_y0=relerr(y0[inner], yr[inner]), err_y1=relerr(y1[inner], yr[inner]),
                                  err_y0b=relerr(y0b[inner], yr[inner])))
    assert d0 == d1
    inner = ~np.repeat(bnd, 3)
    # 3 = k_spmv_dot_tmac (TMA load, TMA store) + k_spmv_fix: all 243 terms per node in the reference's order -- the
    # same bits as the assembled path, also at the nodes whose tile store writes a zero first (k_spmv_fix overwrites it)
    y2, d2 = g.apply_operator(p, op=3, kernel=3)
    assert np.array_equal(y0, y2)
    assert abs(d2 - d0) <= 1e-13 * abs(d0)
    y2b, d2b = g.apply_operator(p, op=3, kernel=3)
    assert np.array_equal(y2, y2b) and d2 == d2b          # deterministic
    r = refpy.RefMicropp(refpy.default_params(**kw))
    A = r.assembly_mat(np.zeros(g.nndim))
    yr = refpy.ell_mvp(*dims, A, p)
    assert relerr(y0[inner], yr[inner]) < 1e-13
