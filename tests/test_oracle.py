"""CPU tests (-m "not gpu") that PIN THE ORACLE: the plain-C restatement oracle/micropp_oracle.c against

  * the golden vectors of the reference's own tests (benchmark-elastic/plastic/damage tables, element-node map,
    mvp known answers, the test_damage history recorded in SURVEY.md appendix B),
  * the committed fixtures tests/golden/*.npz (generated from the compiled reference by tests/golden/make_golden.py),
  * the compiled reference itself (oracle/_ref) on seeded random inputs, whenever it is present.

Integer structures are compared bit-exact; floating point is expected bit-exact too (same operation order, no FMA
contraction) and asserted at <= 1e-13 relative so that a different libm/compiler cannot produce a spurious failure.
"""
from pathlib import Path

import numpy as np
import pytest

from common import CASES, relerr, random_u, random_vars
from oracle import orcpy as O

GOLD = Path(__file__).resolve().parent / "golden"
TIGHT = 1e-13


def load(name):
    return np.load(GOLD / name, allow_pickle=False)


# ------------------------------------------------------------------ golden vectors held by the reference's tests
def test_elem_nodes_reference_golden():
    # test/test_get_elem_nodes.cpp:63-85
    assert list(O.elem_nodes(5, 5, 0, 0, 0)) == [0, 1, 6, 5, 25, 26, 31, 30]
    assert list(O.elem_nodes(5, 5, 3, 3, 3)) == [93, 94, 99, 98, 118, 119, 124, 123]


def test_mvp3_reference_golden():
    # test/test_util_1.cpp:44-56
    a3 = [[10.66, -2.66, 8.3], [10.66, -2.66, 1.9], [-2.66, -10.66, 7.2]]
    y = O.mvp3(a3, [-1.22, 7.1, -11.39])
    assert np.all(np.abs(y - np.array([-126.4282, -53.53220, -154.4488])) < 1e-10)


def test_cols_row_is_neighbour_offset():
    # literal table of src/ell-common.cpp:175-178 == slot of (local node j) - (local node i)
    corner = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    for i in range(8):
        for j in range(8):
            d = [corner[j][q] - corner[i][q] for q in range(3)]
            assert O.cols_row(i, j) == (d[2] + 1) * 9 + (d[1] + 1) * 3 + (d[0] + 1)


GOLD_ELASTIC = [[0.0, 0.0, 0.0], [5.4e4, 1.8e4, 1.8e4], [1.08e5, 3.6e4, 3.6e4], [1.62e5, 5.4e4, 5.4e4],
                [2.16e5, 7.2e4, 7.2e4], [2.7e5, 9.0e4, 9.0e4], [3.24e5, 1.08e5, 1.08e5], [3.78e5, 1.26e5, 1.26e5],
                [4.32e5, 1.44e5, 1.44e5], [4.86e5, 1.62e5, 1.62e5]]  # test/benchmark-elastic.cpp:40-51
GOLD_PLASTIC = [[0.0, 0.0, 0.0], [5.4e4, 1.8e4, 1.8e4], [1.08e5, 3.6e4, 3.6e4],
                [1.57826086961140e+05, 5.60869565194300e+04, 5.60869565194300e+04],
                [1.93043478265488e+05, 8.34782608672561e+04, 8.34782608672561e+04],
                [2.28260869570719e+05, 1.10869565214640e+05, 1.10869565214640e+05],
                [2.63478260875790e+05, 1.38260869562105e+05, 1.38260869562105e+05],
                [2.98695652180861e+05, 1.65652173909570e+05, 1.65652173909570e+05],
                [3.33913043485931e+05, 1.93043478257034e+05, 1.93043478257034e+05],
                [3.69130434791002e+05, 2.20434782604499e+05, 2.20434782604499e+05]]  # test/benchmark-plastic.cpp:40-51
GOLD_DAMAGE = [[0.0, 0.0, 0.0], [5.4e4, 1.8e4, 1.8e4], [1.08e5, 3.6e4, 3.6e4],
               [6.34099396490701e+05, 2.11366465496900e+05, 2.11366465496900e+05],
               [1.13477225575052e+06, 3.78257418583506e+05, 3.78257418583506e+05],
               [1.40477225575052e+06, 4.68257418583505e+05, 4.68257418583505e+05],
               [1.67477225575052e+06, 5.58257418583506e+05, 5.58257418583506e+05],
               [1.94477225575052e+06, 6.48257418583506e+05, 6.48257418583506e+05],
               [2.21477225575052e+06, 7.38257418583506e+05, 7.38257418583506e+05],
               [2.48477225575052e+06, 8.28257418583506e+05, 8.28257418583506e+05]]  # test/benchmark-damage.cpp:40-51


@pytest.mark.parametrize("mat0,gold,tol", [((0, 3e7, 0.25, 0, 0, 0), GOLD_ELASTIC, 1e-10),
                                           ((1, 3e7, 0.25, 1e7, 1e5, 0), GOLD_PLASTIC, 1e-8),
                                           ((2, 3e7, 0.25, 0, 0, 1e5), GOLD_DAMAGE, 1e-8)])
def test_reference_benchmark_tables(mat0, gold, tol):
    # test/benchmark-{elastic,plastic,damage}.cpp: n=2 homogeneous, lin_stress=false, eps_11 = 0.1*t, dt = 0.015
    el = (0, 3e7, 0.25, 0, 0, 0)
    m = O.OrcMicropp(dict(size=(2, 2, 2), type=0, geo_params=(0.1,) * 4, materials=[mat0, el, el], ngp=1,
                          lin_stress=False))
    t = 0.0
    for k in range(10):
        eps = np.zeros(6)
        eps[0] = 0.1 * t
        m.set_strain(0, eps)
        m.homogenize()
        sig = m.get_stress(0)
        assert np.all(np.abs(sig[:3] - np.array(gold[k])) < tol * max(1.0, abs(gold[k][0]) * 1e-6)), (k, sig)
        m.update_vars()
        t += 0.015


def test_reference_test_damage_history():
    # `test_damage 10` observed with the compiled reference: NL flags, CG cost, sigma_22
    el0 = (0, 0.0, 0.0, 0, 0, 0)
    m = O.OrcMicropp(dict(size=(10, 10, 10), type=0, geo_params=(0.1,) * 4, materials=[(2, 1e7, 0.3, 0, 0, 1e5), el0, el0],
                          ngp=1, lin_stress=False, nr_max_its=8))

    def eps_vs_t(t, tf=1.0, emax=0.1):  # test/test_damage.cpp:33-50
        t1, t2, t3 = 0.25 * tf, 0.5 * tf, 0.75 * tf
        if t < t1:
            return 0.5 * emax * t
        if t < t2:
            return 0.5 * emax * (t2 - t)
        if t < t3:
            return emax * (t - t2)
        return emax * (tf - t)

    want_nl = [0, 0, 1, 1, 1, 1, 1, 1, 1, 1]
    want_cost = [0, 90, 114, 93, 52, 40, 81, 68, 40, 42]
    t = 0.0
    sig_hist = []
    for k in range(10):
        eps = np.zeros(6)
        eps[1] = eps_vs_t(t)
        m.set_strain(0, eps)
        m.homogenize()
        assert m.is_non_linear(0) == want_nl[k], k
        assert m.get_cost(0) == want_cost[k], (k, m.get_cost(0))
        assert m.has_converged(0)
        sig_hist.append(m.get_stress(0))
        m.update_vars()
        t += 0.1
    # uniaxial strain in y: sigma_22 of the recorded run (7 significant digits)
    want_s22 = [0, 6.730769e4, 3.019390e5, 3.019704e5, 1.509852e5, None, 3.020312e5, 1.404166e6, 1.404166e6, 7.020829e5]
    for k, w in enumerate(want_s22):
        if w:
            assert abs(sig_hist[k][1] - w) < 2e-6 * abs(w), (k, sig_hist[k])


# ------------------------------------------------------------------ project-defined contract: element colouring
def test_colouring_is_a_valid_8_colouring():
    nx, ny, nz = 6, 5, 7
    seen = {}
    for ez in range(nz - 1):
        for ey in range(ny - 1):
            for ex in range(nx - 1):
                c = O.elem_colour(ex, ey, ez)
                assert 0 <= c < 8
                for nd in O.elem_nodes(nx, ny, ex, ey, ez):
                    assert (c, int(nd)) not in seen, "two elements of one colour share a node"
                    seen[(c, int(nd))] = (ex, ey, ez)
    # union of the colour classes = all element-node incidences
    assert len(seen) == (nx - 1) * (ny - 1) * (nz - 1) * 8


# ------------------------------------------------------------------ committed fixtures (from the compiled reference)
@pytest.mark.parametrize("name,dims", [("ell_cols_3x4x5.npz", (3, 4, 5)), ("ell_cols_2x2x2.npz", (2, 2, 2))])
def test_ell_cols_fixture(name, dims):
    assert np.array_equal(O.ell_cols(*dims), load(name)["cols"])


def test_elem_type_fixture():
    f = load("elem_type_9x11x10.npz")
    for mt in range(5):  # micro-structures restated in the oracle (the others are checked on the product in -m gpu)
        assert np.array_equal(O.elem_types(mt, (0.2, 0.1, 0.1, 0.1), (9, 11, 10)), f[f"type{mt}"]), mt


@pytest.mark.parametrize("case", ["damage_sphere", "plastic_layer", "elastic_sphere", "mic3d_8",
                                  "plastic_layer_yield"])
def test_stage_fixtures(case):
    f = load(f"stages_{case}.npz")
    dims = tuple(int(v) for v in f["dims"])
    p = dict(size=dims, nr_max_its=4, **CASES[case])
    o = O.OrcProblem(p, elem_type=f["elem_type"])
    assert np.array_equal(o.bmat(), f["bmat"])
    assert np.array_equal(o.set_displ_bc(f["eps"], f["u"]), f["u_bc"])
    for tag, vv in (("nov", None), ("v", f["vars"])):
        b, nrm = o.assembly_rhs(f["u"], vv)
        assert relerr(b, f[f"b_{tag}"]) <= TIGHT
        assert abs(nrm - float(f[f"bnorm_{tag}"])) <= TIGHT * nrm
        assert relerr(o.assembly_mat(f["u"], vv), f[f"A_{tag}"]) <= TIGHT
        assert relerr(o.ave_stress(f["u"], vv), f[f"sig_{tag}"]) <= TIGHT
        vn, nl = o.vars_new(f["u"], vv)
        assert nl == bool(f[f"nl_{tag}"])
        assert relerr(vn, f[f"vnew_{tag}"]) <= TIGHT
    assert relerr(O.ell_mvp(*dims, f["cg_A"], f["mvp_x"]), f["mvp_y"]) <= TIGHT
    x, its, err = O.ell_solve_cgpd(*dims, f["cg_A"], f["cg_b"])
    assert its == int(f["cg_its"])
    assert relerr(x, f["cg_x"]) <= 1e-12
    un, st = o.newton(f["nr_eps"], np.zeros(o.nndim))
    assert (st["its"], st["solver_its"], st["converged"]) == (int(f["nr_its"]), int(f["nr_solver_its"]),
                                                             bool(f["nr_conv"]))
    assert relerr(un, f["nr_u"]) <= 1e-12
    assert relerr(o.ave_stress(un), f["nr_sig"]) <= 1e-12


@pytest.mark.parametrize("case", ["damage_sphere", "plastic_layer", "elastic_sphere", "plastic_layer_yield"])
def test_history_fixtures(case):
    f = load(f"history_{case}.npz")
    n, ngp = int(f["n"]), int(f["ngp"])
    m = O.OrcMicropp(dict(size=(n, n, n), ngp=ngp, lin_stress=False, nr_max_its=int(f["nr_max_its"]), **CASES[case]),
                     elem_type=f["elem_type"])
    for k in range(f["eps"].shape[0]):
        for g in range(ngp):
            m.set_strain(g, f["eps"][k, g])
        m.homogenize()
        for g in range(ngp):
            assert m.get_cost(g) == int(f["cost"][k, g]), (k, g)
            assert m.has_converged(g) == bool(f["conv"][k, g])
            assert m.is_non_linear(g) == int(f["nl"][k, g])
            assert relerr(m.get_stress(g), f["sig"][k, g]) <= 1e-11, (k, g)
        m.update_vars()


# ------------------------------------------------------------------ live against the compiled reference (if present)
def test_oracle_vs_compiled_reference_live(refpy):
    rng = np.random.default_rng(99)
    dims = (5, 4, 6)
    for case in ("damage_sphere", "plastic_fibre"):
        p = refpy.default_params(size=dims, calc_ctan_lin=False, **CASES[case])
        r = refpy.RefMicropp(p)
        o = O.OrcProblem(dict(p), elem_type=r.elem_type())
        mt = [p["materials"][t][0] for t in r.elem_type()]
        u = random_u(r.nndim, int(rng.integers(1 << 30)), 1e-2)
        v = random_vars(r.nelem, mt, int(rng.integers(1 << 30)))
        assert np.array_equal(o.elem_type(), O.elem_types(p["type"], p["geo_params"], dims))
        assert relerr(o.assembly_rhs(u, v)[0], r.assembly_rhs(u, v)[0]) <= TIGHT
        assert relerr(o.assembly_mat(u, v), r.assembly_mat(u, v)) <= TIGHT
        assert relerr(o.ave_stress(u, v), r.ave_stress(u, v)) <= TIGHT
        r.close()
    m = (2, 1e7, 0.3, 0, 0, 1e5)
    for _ in range(50):
        eps = rng.uniform(-1, 1, 6) * 10.0 ** rng.uniform(-5, -1)
        assert np.array_equal(O.mat_stress(m, eps), refpy.mat_stress(m, eps))
        assert np.array_equal(O.mat_ctan(m, eps), refpy.mat_ctan(m, eps))


# ------------------------------------------------------------------ full-size fixtures: inputs are bench.py's own
@pytest.mark.parametrize("workload", ["damage50", "plastic40"])
def test_full_size_fixture_inputs_are_the_bench_load_path(workload):
    """tests/golden/bench_<workload>_2gp.npz (reference run at the bench size, make_golden_full.py) was generated from
    the very strains bench.py feeds the product, and its Gauss points end non-linear."""
    import bench
    f = load(f"bench_{workload}_2gp.npz")
    wl = bench.WORKLOADS[workload]
    assert int(f["n"]) == wl["n"]
    gps = [int(g) for g in f["gps"]]
    for k in range(f["eps"].shape[0]):
        assert np.array_equal(f["eps"][k], bench.strains_for(workload, wl["ngp"], 0, k)[gps])
    assert f["nl"][-1].all() and np.all(np.isfinite(f["sig"]))
    assert f["eps"].shape[0] == wl["prep_steps"] + 1          # the fixture covers the step bench.py times


def test_cg_residual_history_is_the_plain_solver():
    """orc_ell_solve_cgpd_hist is orc_ell_solve_cgpd (src/ell.cpp:66-122) + a record of the |z| its loop-head test sees:
    same iterate, same count; hist[0] = |z_0|, the last recorded value is the first one under the tolerance."""
    p = O.OrcProblem(dict(size=(6, 5, 7), lin_stress=False, **CASES["elastic_sphere"]))
    u = p.set_displ_bc(np.array([1e-3, -2e-4, 3e-4, 5e-4, 0.0, -1e-4]))
    b, _ = p.assembly_rhs(u)
    vals = p.assembly_mat(u)
    x0, its0, _ = O.ell_solve_cgpd(6, 5, 7, vals, b)
    x1, its1, hist = O.ell_solve_cgpd_hist(6, 5, 7, vals, b, 1000)
    assert its0 == its1 and np.array_equal(x0, x1) and len(hist) == its1 + 1
    assert np.all(hist[:-1] >= hist[0] * 1e-5) and hist[-1] < hist[0] * 1e-5
    _, _, short = O.ell_solve_cgpd_hist(6, 5, 7, vals, b, 4)
    assert np.array_equal(short, hist[:4])
