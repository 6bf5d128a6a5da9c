"""-m gpu: the implicit operator of all-elastic RVEs (no per-slot ELL matrix; DPCG applies the table of distinct
row blocks to 8 right-hand sides per thread) against (a) the explicit per-slot-matrix path of the same library,
which must agree BIT FOR BIT (same row values, same FMA order, same reduction tree), and (b) the reference."""
import os

import numpy as np
import pytest

from common import CASES, relerr

pytestmark = pytest.mark.gpu


def run(mod_cls, params, eps, implicit=None):
    old = os.environ.get("MICROPP_IMPLICIT")
    if implicit is not None:
        os.environ["MICROPP_IMPLICIT"] = "1" if implicit else "0"
    try:
        g = mod_cls(params)
    finally:
        if implicit is not None:
            if old is None:
                del os.environ["MICROPP_IMPLICIT"]
            else:
                os.environ["MICROPP_IMPLICIT"] = old
    ngp = eps.shape[0]
    for gp in range(ngp):
        g.set_strain(gp, eps[gp])
    g.homogenize()
    sig = np.array([g.get_stress(gp) for gp in range(ngp)])
    cost = [g.get_cost(gp) for gp in range(ngp)]
    conv = [g.has_converged(gp) for gp in range(ngp)]
    return g, sig, cost, conv


@pytest.mark.parametrize("case,dims,ngp", [("elastic_sphere", (12, 12, 12), 19), ("elastic_sphere", (9, 11, 10), 8),
                                            ("elastic_sphere", (14, 9, 8), 3)])
def test_implicit_equals_explicit_bitwise(mpp, case, dims, ngp):
    rng = np.random.default_rng(42)
    eps = rng.uniform(-1e-3, 1e-3, (ngp, 6))
    kw = dict(size=dims, ngp=ngp, lin_stress=False, calc_ctan_lin=False, **CASES[case])
    gi, si, ci, vi = run(mpp.Micropp3, mpp.default_params(**kw), eps, implicit=True)
    ge, se, ce, ve = run(mpp.Micropp3, mpp.default_params(**kw), eps, implicit=False)
    assert gi.implicit_rows() > 3          # sphere interface: more row blocks than materials
    assert ge.implicit_rows() == 0
    assert ci == ce and vi == ve
    assert np.array_equal(si, se)
    for gp in (0, ngp - 1):
        assert np.array_equal(gi.get_u(gp), ge.get_u(gp))


def test_implicit_homogeneous_single_row(mpp):
    kw = dict(size=(6, 6, 6), ngp=1, type=0, materials=[(0, 3.0e7, 0.25, 0, 0, 0)] * 3, lin_stress=False,
              calc_ctan_lin=False)
    g, sig, cost, conv = run(mpp.Micropp3, mpp.default_params(**kw), np.array([[1e-3, 0, 0, 0, 0, 0]]), implicit=True)
    assert g.implicit_rows() == 1
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
