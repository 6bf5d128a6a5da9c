"""-m gpu: the CUDA path (through the C ABI) against the COMMITTED golden fixtures of tests/golden/ (generated from the
compiled reference by tests/golden/make_golden.py) and, at BASELINE.json's full RVE sizes, through size-independent
properties (exact linearity in powers of two, rigid-body null space of the assembled operator, GP independence,
the plain-C oracle's SpMV on the product's own matrix).

Tolerances: integer structures bit-exact; FE stages on identical inputs <= 1e-12 relative (1e-8 for the forward-
difference Jacobians, which amplify rounding by 1/D_EPS_CTAN = 1e8); homogenized stress 1e-8 relative; Newton / CG
iteration counts within +-1 per Newton step (BASELINE.json north_star).
"""
from pathlib import Path

import numpy as np
import pytest

from common import CASES, relerr

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def load(name):
    return np.load(GOLD / name, allow_pickle=False)


@pytest.mark.parametrize("name,dims", [("ell_cols_3x4x5.npz", (3, 4, 5)), ("ell_cols_2x2x2.npz", (2, 2, 2))])
def test_ell_cols_fixture(mpp, name, dims):
    assert np.array_equal(mpp.ell_cols(*dims), load(name)["cols"])


def test_elem_type_all_13_microstructures(mpp):
    f = load("elem_type_9x11x10.npz")
    for mt in range(13):
        m = mpp.Micropp3(mpp.default_params(size=(9, 11, 10), type=mt, geo_params=(0.2, 0.1, 0.1, 0.1),
                                            calc_ctan_lin=False))
        assert np.array_equal(m.elem_type(), f[f"type{mt}"]), mt
        m.close()


def test_colouring_contract(mpp):
    from oracle import orcpy as O
    for ez in range(4):
        for ey in range(5):
            for ex in range(3):
                assert mpp.elem_colour(ex, ey, ez) == O.elem_colour(ex, ey, ez) == (ex & 1) + 2 * (ey & 1) + 4 * (ez & 1)


@pytest.mark.parametrize("case", ["damage_sphere", "plastic_layer", "elastic_sphere", "mic3d_8",
                                  "plastic_layer_yield"])
def test_stage_fixtures(mpp, case):
    f = load(f"stages_{case}.npz")
    dims = tuple(int(v) for v in f["dims"])
    g = mpp.Micropp3(mpp.default_params(size=dims, calc_ctan_lin=False, **CASES[case]))
    assert np.array_equal(g.elem_type(), f["elem_type"])
    assert np.array_equal(g.bmat(), f["bmat"])
    assert np.array_equal(g.set_displ_bc(f["eps"], f["u"]), f["u_bc"])
    jtol = 1e-12 if case == "elastic_sphere" else 1e-8
    for tag, vv in (("nov", None), ("v", f["vars"])):
        b, nrm = g.assembly_rhs(f["u"], vv)
        assert relerr(b, f[f"b_{tag}"]) < 1e-12
        assert abs(nrm - float(f[f"bnorm_{tag}"])) <= 1e-12 * nrm
        assert relerr(g.assembly_mat(f["u"], vv), f[f"A_{tag}"]) < jtol
        assert relerr(g.ave_stress(f["u"], vv), f[f"sig_{tag}"]) < 1e-12
        vn, nl = g.vars_new(f["u"], vv)
        assert nl == bool(f[f"nl_{tag}"])
        assert relerr(vn, f[f"vnew_{tag}"]) < 1e-13
    assert relerr(mpp.ell_mvp(*dims, f["cg_A"], f["mvp_x"]), f["mvp_y"]) < 1e-13
    x, its, err = mpp.ell_solve_cgpd(*dims, f["cg_A"], f["cg_b"])
    assert abs(its - int(f["cg_its"])) <= 1
    assert relerr(x, f["cg_x"]) < 1e-8
    g2 = mpp.Micropp3(mpp.default_params(size=dims, calc_ctan_lin=False, **CASES[case]))
    un, st = g2.newton(f["nr_eps"], np.zeros(g2.nndim))
    assert st["converged"] == bool(f["nr_conv"])
    assert abs(st["its"] - int(f["nr_its"])) <= 1
    assert abs(st["solver_its"] - int(f["nr_solver_its"])) <= max(1, int(f["nr_its"]))
    assert relerr(g2.ave_stress(un), f["nr_sig"]) < 1e-8


@pytest.mark.parametrize("case", ["damage_sphere", "plastic_layer", "elastic_sphere", "plastic_layer_yield"])
def test_history_fixtures(mpp, case):
    f = load(f"history_{case}.npz")
    n, ngp, nr = int(f["n"]), int(f["ngp"]), int(f["nr_max_its"])
    m = mpp.Micropp3(mpp.default_params(size=(n, n, n), ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=nr,
                                        **CASES[case]))
    for k in range(f["eps"].shape[0]):
        m.set_strains(f["eps"][k])
        m.homogenize()
        sig = m.get_stresses()
        for gp in range(ngp):
            assert m.is_non_linear(gp) == int(f["nl"][k, gp]), (k, gp)
            assert m.has_converged(gp) == bool(f["conv"][k, gp]), (k, gp)
            assert abs(m.get_cost(gp) - int(f["cost"][k, gp])) <= nr, (k, gp, m.get_cost(gp), int(f["cost"][k, gp]))
            assert relerr(sig[gp], f["sig"][k, gp]) < 1e-8, (k, gp)
        m.update_vars()
    if case in ("damage_sphere", "plastic_layer_yield"):
        assert f["nl"][-1].all()  # the fixture (and therefore the product) went through the non-linear branch


def test_ctan_lin_fixture(mpp):
    m = mpp.Micropp3(mpp.default_params(size=(6, 6, 6), **CASES["damage_sphere"]))
    assert relerr(m.ctan_lin(), load("ctan_lin_damage_sphere_6.npz")["ctan_lin"]) < 1e-8


# ------------------------------------------------------------------------- BASELINE sizes: size-independent properties
@pytest.mark.parametrize("n", [30, 50])
def test_full_size_elastic_linearity_and_independence(mpp, n):
    # configs[1] RVE (and its 50^3 sibling): scaling the strain by a power of two scales every intermediate of
    # Newton/DPCG exactly, so stresses must scale bit-exactly and iteration counts must not change; identical
    # strains on different GPs must give bit-identical results (deterministic reductions).
    ngp = 4
    m = mpp.Micropp3(mpp.default_params(size=(n, n, n), ngp=ngp, lin_stress=False, calc_ctan_lin=False,
                                        **CASES["elastic_sphere"]))
    e = np.random.default_rng(1234).uniform(-1e-3, 1e-3, 6)
    m.set_strains(np.array([e, 2 * e, 0.25 * e, e]))
    m.homogenize()
    s = m.get_stresses()
    c = [m.get_cost(g) for g in range(ngp)]
    assert np.array_equal(s[1], 2 * s[0]) and np.array_equal(s[2], 0.25 * s[0]) and np.array_equal(s[3], s[0])
    assert len(set(c)) == 1 and c[0] > 20
    assert all(m.has_converged(g) for g in range(ngp))
    # symmetry of the homogenized response: sigma(e) . e > 0 for an elastic RVE
    assert float(np.dot(s[0], e)) > 0


@pytest.mark.parametrize("case,n", [("elastic_sphere", 30), ("damage_sphere", 50)])
def test_full_size_operator_null_space_and_spmv(mpp, case, n):
    # assembled Jacobian at BASELINE sizes: interior rows annihilate rigid translations (B t = 0), boundary rows are
    # identity rows; the CUDA SpMV on that matrix equals the oracle's CPU SpMV (src/ell.cpp:35-44 restated).
    from oracle import orcpy as O
    g = mpp.Micropp3(mpp.default_params(size=(n, n, n), calc_ctan_lin=False, **CASES[case]))
    rng = np.random.default_rng(7)
    u = g.set_displ_bc(np.array([0.01, -0.004, 0.002, 0.006, -0.003, 0.001]), rng.uniform(-1e-3, 1e-3, g.nndim))
    A = g.assembly_mat(u)
    t3 = np.array([1.0, -2.0, 0.5])
    y = (A.reshape(g.nn, 3, 27, 3) * t3.reshape(1, 1, 1, 3)).sum(axis=(2, 3))  # every neighbour carries t3
    idx = np.arange(g.nn)
    i, j, k = idx % n, (idx // n) % n, idx // (n * n)
    bnd = (i == 0) | (i == n - 1) | (j == 0) | (j == n - 1) | (k == 0) | (k == n - 1)
    scale = np.abs(A).max()
    assert np.abs(y[~bnd]).max() < 1e-9 * scale
    assert np.array_equal(y[bnd], np.tile(np.array([1.0, -2.0, 0.5]), (int(bnd.sum()), 1)))
    x = rng.uniform(-1, 1, g.nndim)
    assert relerr(mpp.ell_mvp(n, n, n, A, x), O.ell_mvp(n, n, n, A, x)) < 1e-13


def test_full_size_damage_gp_independence(mpp):
    # configs[2] RVE: two GPs with the same strain history stay bit-identical through the non-linear regime
    n, ngp = 50, 3
    m = mpp.Micropp3(mpp.default_params(size=(n, n, n), ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=12,
                                        **CASES["damage_sphere"]))
    for k in (3, 7):
        e = np.zeros((ngp, 6))
        e[:, 0] = np.array([1.3, 0.7, 1.3]) * 0.1 * 0.015 * k
        m.set_strains(e)
        m.homogenize()
        s = m.get_stresses()
        assert np.array_equal(s[0], s[2]) and m.get_cost(0) == m.get_cost(2)
        assert np.all(np.isfinite(s))
        m.update_vars()
    assert m.is_non_linear(0) == 1 and m.is_non_linear(2) == 1
