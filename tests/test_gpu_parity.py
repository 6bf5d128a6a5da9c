"""-m gpu: every stage of the hot path, called THROUGH THE C-ABI of libmicropp_b200.so, against the
reference CPU implementation (oracle/_ref) on identical seeded inputs.

Tolerances (BASELINE.json north_star): index structures bit-exact; stress / tangent 1e-8 relative;
Newton and CG iteration counts within +-1.  Stage tests with identical inputs are held much tighter.
"""
import numpy as np
import pytest

from common import CASES, relerr, random_u, random_vars

pytestmark = pytest.mark.gpu


def mk(mod, case, n, **kw):
    size = n if isinstance(n, tuple) else (n, n, n)
    p = mod.default_params(size=size, **CASES[case])
    p.update(kw)
    return p


def pair(mpp, refpy, case, n, **kw):
    g = mpp.Micropp3(mk(mpp, case, n, **kw))
    r = refpy.RefMicropp(mk(refpy, case, n, **kw))
    return g, r


def mat_types(ref):
    et = ref.elem_type()
    mtypes = [m[0] for m in ref.p["materials"]]
    return [mtypes[t] for t in et]


# ---------------------------------------------------------------- integer structures: bit-exact
@pytest.mark.parametrize("dims", [(2, 2, 2), (3, 4, 5), (6, 5, 4), (9, 9, 9)])
def test_ell_cols_bit_exact(mpp, refpy, dims):
    assert np.array_equal(mpp.ell_cols(*dims), refpy.ell_cols(*dims))


def test_elem_nodes_golden(mpp, refpy):
    # test/test_get_elem_nodes.cpp:63-85
    assert list(mpp.elem_nodes(5, 5, 0, 0, 0)) == [0, 1, 6, 5, 25, 26, 31, 30]
    for (ex, ey, ez) in [(0, 0, 0), (3, 3, 3), (1, 2, 3), (0, 3, 1)]:
        assert np.array_equal(mpp.elem_nodes(5, 5, ex, ey, ez), refpy.elem_nodes(5, 5, ex, ey, ez))


@pytest.mark.parametrize("case", ["elastic_sphere", "plastic_layer", "plastic_fibre", "mic3d_8"])
def test_elem_type_bit_exact(mpp, refpy, case):
    g, r = pair(mpp, refpy, case, (7, 9, 8), calc_ctan_lin=False)
    assert np.array_equal(g.elem_type(), r.elem_type())
    assert np.array_equal(g.bmat(), r.bmat())


# ---------------------------------------------------------------- FE stages on identical inputs
@pytest.mark.parametrize("dims", [(5, 5, 5), (4, 6, 7)])
def test_set_displ_bc(mpp, refpy, dims):
    g, r = pair(mpp, refpy, "elastic_sphere", dims, calc_ctan_lin=False)
    eps = np.array([1.0, 2.0, 3.0, 1.0, 1.0, 1.0]) * 1e-3
    u0 = random_u(g.nndim, 1)
    ug, ur = g.set_displ_bc(eps, u0), r.set_displ_bc(eps, u0)
    assert np.array_equal(ug, ur)  # non-contracted arithmetic, same face-overwrite order


@pytest.mark.parametrize("case,with_vars", [("elastic_sphere", False), ("damage_sphere", False),
                                            ("damage_sphere", True), ("plastic_layer", False),
                                            ("plastic_layer", True), ("mic3d_8", True),
                                            ("plastic_layer_yield", False), ("plastic_layer_yield", True),
                                            ("plastic_fibre_yield", True)])
def test_assembly_rhs(mpp, refpy, case, with_vars):
    g, r = pair(mpp, refpy, case, (7, 6, 8), calc_ctan_lin=False)
    u = random_u(g.nndim, 2, 5e-3)
    v = random_vars(r.nelem, mat_types(r), 3) if with_vars else None
    bg, ng = g.assembly_rhs(u, v)
    br, nr = r.assembly_rhs(u, v)
    assert relerr(bg, br) < 1e-12
    assert abs(ng - nr) <= 1e-12 * nr


@pytest.mark.parametrize("case,with_vars", [("elastic_sphere", False), ("damage_sphere", False),
                                            ("damage_sphere", True), ("plastic_layer", True), ("mic3d_8", True),
                                            ("plastic_layer_yield", False), ("plastic_layer_yield", True),
                                            ("plastic_fibre_yield", True)])
def test_assembly_mat(mpp, refpy, case, with_vars):
    g, r = pair(mpp, refpy, case, (6, 7, 5), calc_ctan_lin=False)
    u = random_u(g.nndim, 4, 5e-3)
    v = random_vars(r.nelem, mat_types(r), 5) if with_vars else None
    Ag, Ar = g.assembly_mat(u, v), r.assembly_mat(u, v)
    # forward-difference tangents amplify rounding by 1e8: identical inputs still agree to ~1e-9
    tol = 1e-12 if case == "elastic_sphere" else 1e-8
    assert relerr(Ag, Ar) < tol


def test_ave_stress_and_vars_new(mpp, refpy):
    for case in ("damage_sphere", "plastic_layer", "plastic_layer_yield"):
        g, r = pair(mpp, refpy, case, (6, 6, 6), calc_ctan_lin=False)
        u = random_u(g.nndim, 6, 2e-2)
        v = random_vars(r.nelem, mat_types(r), 7)
        for vv in (None, v):
            assert relerr(g.ave_stress(u, vv), r.ave_stress(u, vv)) < 1e-12
            vg, fg = g.vars_new(u, vv)
            vr, fr = r.vars_new(u, vv)
            assert fg == fr
            assert relerr(vg, vr) < 1e-13


# ---------------------------------------------------------------- ELL SpMV + DPCG
def test_ell_mvp_and_cg(mpp, refpy):
    n = (8, 7, 9)
    g, r = pair(mpp, refpy, "elastic_sphere", n, calc_ctan_lin=False)
    eps = np.array([1.0, 2.0, 3.0, 1.0, 1.0, 1.0])  # test/test_cg.cpp:33
    u = r.set_displ_bc(eps)
    A = r.assembly_mat(u)
    b, _ = r.assembly_rhs(u)
    x = random_u(g.nndim, 8, 1.0)
    yg, yr = mpp.ell_mvp(*n, A, x), refpy.ell_mvp(*n, A, x)
    assert relerr(yg, yr) < 1e-13
    xg, ig, eg = mpp.ell_solve_cgpd(*n, A, b)
    xr, ir, er = refpy.ell_solve_cgpd(*n, A, b)
    assert abs(ig - ir) <= 1
    assert relerr(xg, xr) < 1e-8
    assert abs(eg - er) <= 1e-6 * abs(er)


@pytest.mark.parametrize("case", ["elastic_sphere", "damage_sphere", "plastic_layer", "plastic_layer_yield"])
def test_newton(mpp, refpy, case):
    g, r = pair(mpp, refpy, case, 9, calc_ctan_lin=False, nr_max_its=6)
    eps = np.array([0.01, -0.004, 0.002, 0.006, -0.003, 0.001]) * (1.0 if case != "elastic_sphere" else 0.1)
    u0 = np.zeros(g.nndim)
    ug, sg = g.newton(eps, u0)
    ur, sr = r.newton(eps, u0)
    assert sg["converged"] == sr["converged"]
    assert abs(sg["its"] - sr["its"]) <= 1
    assert abs(sg["solver_its"] - sr["solver_its"]) <= max(1, sr["its"])
    assert relerr(ug, ur) < 1e-7
    assert relerr(g.ave_stress(ug), r.ave_stress(ur)) < 1e-8


# ---------------------------------------------------------------- golden tables of the reference's own tests
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
def test_reference_golden_tables(mpp, mat0, gold, tol):
    el = (0, 3e7, 0.25, 0, 0, 0)
    m = mpp.Micropp3(size=(2, 2, 2), type=0, materials=[mat0, el, el], lin_stress=False)
    t = 0.0
    for k in range(10):
        eps = np.zeros(6)
        eps[0] = 0.1 * t
        m.set_strain(0, eps)
        m.homogenize()
        sig = m.get_stress(0)
        assert np.all(np.abs(sig[:3] - np.array(gold[k])) < tol * max(1.0, abs(gold[k][0]) * 1e-6)), (k, sig)
        assert np.all(np.abs(sig[3:]) < tol)
        m.update_vars()
        t += 0.015


# ---------------------------------------------------------------- full homogenize() histories vs the reference
def run_history(obj, strains_per_step):
    out = []
    ngp = obj.ngp
    for eps_all in strains_per_step:
        for gp in range(ngp):
            obj.set_strain(gp, eps_all[gp])
        obj.homogenize()
        out.append(dict(sig=np.array([obj.get_stress(gp) for gp in range(ngp)]),
                        ctan=np.array([obj.get_ctan(gp) for gp in range(ngp)]),
                        cost=[obj.get_cost(gp) for gp in range(ngp)],
                        conv=[obj.has_converged(gp) for gp in range(ngp)],
                        sub=[obj.has_subiterated(gp) for gp in range(ngp)],
                        nl=[obj.is_non_linear(gp) for gp in range(ngp)]))
        obj.update_vars()
    return out


def compare_histories(hg, hr, sig_tol=1e-8, ctan_tol=None, newton_budget=12):
    for k, (a, b) in enumerate(zip(hg, hr)):
        assert a["nl"] == b["nl"], (k, a["nl"], b["nl"])
        assert a["conv"] == b["conv"], (k, a["conv"], b["conv"])
        assert a["sub"] == b["sub"], k
        for gp in range(len(a["cost"])):
            # +-1 CG iteration per Newton step
            assert abs(a["cost"][gp] - b["cost"][gp]) <= newton_budget, (k, gp, a["cost"], b["cost"])
            e = relerr(a["sig"][gp], b["sig"][gp])
            assert e < sig_tol, (k, gp, e, a["sig"][gp], b["sig"][gp])
            if ctan_tol is not None:
                assert relerr(a["ctan"][gp], b["ctan"][gp]) < ctan_tol, (k, gp)


def load_path(ngp, steps, seed, comp=0, eps_max=0.1, dt=0.015):
    rng = np.random.default_rng(seed)
    scale = rng.uniform(0.5, 1.5, ngp)
    hist = []
    for k in range(steps):
        e = np.zeros((ngp, 6))
        e[:, comp] = scale * eps_max * dt * k
        hist.append(e)
    return hist


def test_homogenize_elastic_batch(mpp, refpy):
    # BASELINE config 2 shrunk: many GPs, random strains, one step
    ngp = 7
    g, r = pair(mpp, refpy, "elastic_sphere", 10, ngp=ngp, lin_stress=False, calc_ctan_lin=False)
    rng = np.random.default_rng(1234)
    eps = rng.uniform(-1e-3, 1e-3, (ngp, 6))
    hg, hr = run_history(g, [eps]), run_history(r, [eps])
    compare_histories(hg, hr)


def test_homogenize_damage_history(mpp, refpy):
    # BASELINE config 3 shrunk (test/benchmark-mic-2.cpp load path; nr_max_its of test/benchmark-sc2019.cpp:83)
    ngp = 4
    kw = dict(ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=12)
    g, r = pair(mpp, refpy, "damage_sphere", 8, **kw)
    path = load_path(ngp, 8, 1234)
    compare_histories(run_history(g, path), run_history(r, path), newton_budget=14)


def test_homogenize_plastic_history(mpp, refpy):
    # BASELINE config 4 shrunk (test/test3d_4.cpp:35,65-67,80-85): load then unload in eps_22... component 1
    ngp = 3
    kw = dict(ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=8)
    g, r = pair(mpp, refpy, "plastic_layer", 8, **kw)
    rng = np.random.default_rng(7)
    scale = rng.uniform(0.5, 1.5, ngp)
    path, e = [], np.zeros((ngp, 6))
    for k in range(10):
        e = e.copy()
        e[:, 1] += (0.01 if k < 6 else -0.01) * scale
        path.append(e)
    compare_histories(run_history(g, path), run_history(r, path), newton_budget=10)


@pytest.mark.parametrize("case,comp", [("plastic_layer_yield", 2), ("plastic_fibre_yield", 1)])
def test_homogenize_plastic_yield_load_unload(mpp, refpy, case, comp):
    """J2 plasticity that REALLY yields (E = 3e7, Sy = 1e5: the reference's golden plastic material and bench.py's
    plastic40 material): heterogeneous RVE, load for 7 steps then unload (test/test3d_4.cpp:80-85), against the compiled
    reference -- src/material.cpp:111-186 (return mapping, evolute) and :49-63 (forward-difference tangent in the
    yielded state) through k_elem_rhs / k_elem_ctan / k_asm_mat_general / k_vars_new and a DPCG solve per Newton step."""
    ngp = 3
    kw = dict(ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=12)
    g, r = pair(mpp, refpy, case, 9, **kw)
    scale = np.random.default_rng(17).uniform(0.5, 1.5, ngp)
    path, e = [], np.zeros((ngp, 6))
    for k in range(12):
        e = e.copy()
        e[:, comp] += (0.0015 if k < 7 else -0.0015) * scale
        path.append(e)
    hg, hr = run_history(g, path), run_history(r, path)
    assert all(hr[-1]["nl"]) and any(hr[3]["nl"])          # the reference itself went plastic on the way up
    assert len({c for h in hr for c in h["cost"]}) > 3     # ... and the work per step changed with the plastic state
    compare_histories(hg, hr, newton_budget=12)


def test_homogenize_defaults_lin_stress_and_ctan_lin(mpp, refpy):
    # defaults of the C/Fortran entry: lin_stress=true, calc_ctan_lin=true
    g, r = pair(mpp, refpy, "damage_sphere", 6, ngp=2)
    assert relerr(g.ctan_lin(), r.ctan_lin()) < 1e-8
    path = load_path(2, 5, 3)
    compare_histories(run_history(g, path), run_history(r, path), ctan_tol=1e-8)


def test_homogenize_fe_full_and_subiterations(mpp, refpy):
    ngp = 3
    cpl = [mpp.FE_FULL, mpp.FE_ONE_WAY, mpp.FE_FULL]
    kw = dict(ngp=ngp, coupling=cpl, lin_stress=False, calc_ctan_lin=True, nr_max_its=2, subiterations=True,
              nsubiterations=3)
    g, r = pair(mpp, refpy, "homog_damage", 5, **kw)
    path = load_path(ngp, 6, 11)
    hg, hr = run_history(g, path), run_history(r, path)
    # the perturbation tangent divides O(1e-5)-accurate stresses by 1e-8: only its bookkeeping is comparable
    compare_histories(hg, hr, newton_budget=40)


def test_fe_full_ctan_values(mpp, refpy):
    """FE_FULL homogenized tangent (src/homogenize.cpp:252-276) compared VALUE by value.  ctan[:, i] =
    (sigma(eps + 1e-8 e_i) - sigma(eps)) / 1e-8 amplifies any difference in sigma by 1e8, so both sides solve the
    Newton systems to the limit (nr_rel_tol = 1e-10, nr_max_its = 20; DPCG gains 1e-5 per Newton step): what is left
    is rounding (different summation orders), about 1e-16 |sigma| / 1e-8."""
    ngp = 3
    cpl = [mpp.FE_FULL, mpp.FE_FULL, mpp.FE_ONE_WAY]
    kw = dict(ngp=ngp, coupling=cpl, lin_stress=False, calc_ctan_lin=True, nr_max_its=20, nr_rel_tol=1e-10)
    for case, n in (("damage_sphere", 7), ("plastic_layer_yield", 6)):
        g, r = pair(mpp, refpy, case, n, **kw)
        path = load_path(ngp, 7, 23, comp=(2 if case.startswith("plastic") else 0), eps_max=0.1)
        hg, hr = run_history(g, path), run_history(r, path)
        assert hr[-1]["nl"][0] and hr[-1]["nl"][1]
        worst = 0.0
        for k, (a, b) in enumerate(zip(hg, hr)):
            assert a["nl"] == b["nl"] and a["conv"] == b["conv"], k
            for gp in range(ngp):
                assert relerr(a["sig"][gp], b["sig"][gp]) < 1e-10, (case, k, gp)
                worst = max(worst, relerr(a["ctan"][gp], b["ctan"][gp]))
        print(f"FE_FULL ctan {case}: worst relative difference to the reference {worst:.3e}")
        assert worst < 1e-8, (case, worst)   # measured on the B200: 1.2e-9 (damage), 1.0e-9 (plastic)
        # the tangent really is the non-linear one (differs from the linear ctan of the constructor)
        assert relerr(hg[-1]["ctan"][0], g.ctan_lin()) > 1e-3


@pytest.mark.parametrize("case,its_A0", [("damage_sphere", 1), ("damage_sphere", 2), ("plastic_layer_yield", 1)])
def test_use_A0_non_elastic(mpp, refpy, case, its_A0):
    """use_A0 (src/solve.cpp:56-66, test/test_A0.cpp): the first its_with_A0 Newton iterations of every solve use the
    shared LINEAR Jacobian assembled once in the constructor (OP_SHARED on the device) -- on a damage / plastic RVE,
    where that matrix differs from the current tangent, so iteration counts and paths really depend on it."""
    ngp = 3
    kw = dict(ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=12, use_A0=True, its_with_A0=its_A0)
    g, r = pair(mpp, refpy, case, 8, **kw)
    g0, _ = pair(mpp, refpy, case, 8, **dict(kw, use_A0=False))
    path = load_path(ngp, 8, 1234, comp=(2 if case.startswith("plastic") else 0))
    hg, hr, h0 = run_history(g, path), run_history(r, path), run_history(g0, path)
    assert all(hr[-1]["nl"])
    compare_histories(hg, hr, newton_budget=12)
    # A0 changed the work (otherwise this test would not test anything)
    assert [h["cost"] for h in hg] != [h["cost"] for h in h0]


def test_write_log_file(mpp, refpy, tmp_path, monkeypatch):
    """write_log (src/output.cpp:265-282, file opened in the constructor src/micropp.cpp:168-175): same file name, same
    bytes as the reference's log for the same history."""
    kw = dict(ngp=3, lin_stress=False, calc_ctan_lin=False, nr_max_its=12, write_log=True, mpi_rank=0)
    path = load_path(3, 6, 4)
    texts = []
    for name, mod in (("ours", mpp), ("ref", refpy)):
        d = tmp_path / name
        d.mkdir()
        monkeypatch.chdir(d)
        o = (mod.Micropp3 if mod is mpp else mod.RefMicropp)(mk(mod, "damage_sphere", 6, **kw))
        h = run_history(o, path)
        o.close()
        texts.append(((d / "micropp-profiling-0.log").read_text(), h))
    (to, ho), (tr, hr) = texts
    assert to.splitlines()[0] == tr.splitlines()[0] == "#<gp_id>  <non-linear>  <cost>  <converged>"
    assert len(to.splitlines()) == len(tr.splitlines()) == 1 + 6 * 4
    if [h["cost"] for h in ho] == [h["cost"] for h in hr]:
        assert to == tr
    else:  # a CG count may differ by one: every other field must still be identical
        for lo, lr in zip(to.splitlines(), tr.splitlines()):
            fo, fr = lo.split(), lr.split()
            assert fo[:2] == fr[:2] and fo[3:] == fr[3:], (lo, lr)


@pytest.mark.parametrize("case", ["plastic_fibre", "plastic_fibre_yield"])
def test_gp_independence(mpp, case):
    # test/test3d_4.cpp:109-115: identical strains => identical results on every GP
    m = mpp.Micropp3(mpp.default_params(size=(5, 5, 5), ngp=5, lin_stress=False, calc_ctan_lin=False,
                                        **CASES[case]))
    e = np.zeros(6)
    for k in range(6):
        e[1] += 0.01 if case == "plastic_fibre" else 0.002
        for gp in range(5):
            m.set_strain(gp, e)
        m.homogenize()
        s = np.array([m.get_stress(gp) for gp in range(5)])
        assert np.all(np.abs(s - s[0]) == 0.0)  # deterministic reductions: bit-identical across slots
        m.update_vars()


def test_c_api_constructor_and_linear(mpp, refpy):
    # micropp3_new (include/micropp_c.h): lin_stress / calc_ctan_lin defaults, subiterations forced on
    cpl = [mpp.FE_ONE_WAY, mpp.FE_LINEAR]
    p = mk(mpp, "damage_sphere", 5, ngp=2, coupling=cpl, nsubiterations=4)
    g = mpp.Micropp3(p, c_api=True)
    pr = mk(refpy, "damage_sphere", 5, ngp=2, coupling=cpl, nsubiterations=4, subiterations=True)
    r = refpy.RefMicropp(pr)
    path = load_path(2, 4, 5)
    compare_histories(run_history(g, path), run_history(r, path), ctan_tol=1e-8)


def test_restart_roundtrip(mpp, refpy, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    kw = dict(ngp=2, lin_stress=False, calc_ctan_lin=False, nr_max_its=10)
    g, r = pair(mpp, refpy, "homog_damage", 5, **kw)
    path = load_path(2, 12, 9)
    run_history(g, path[:10])
    run_history(r, path[:10])
    assert g.get_non_linear_gps() == 2 and r.get_non_linear_gps() == 2
    g.write_restart(7)
    ours = (tmp_path / "micropp-restart-0-7.bin").read_bytes()
    r.write_restart(8)
    theirs = (tmp_path / "micropp-restart-0-8.bin").read_bytes()
    assert len(ours) == len(theirs)  # same on-disk format (src/output.cpp:217-262)
    a = np.frombuffer(ours[1:1 + 8 * g.nvars], dtype=np.float64)
    b = np.frombuffer(theirs[1:1 + 8 * g.nvars], dtype=np.float64)
    assert relerr(a, b) < 1e-7
    # a fresh object restarted from OUR file continues exactly like the original
    g2 = mpp.Micropp3(mk(mpp, "homog_damage", 5, **kw))
    g2.read_restart(7)
    # ... and one restarted from the REFERENCE's file continues like the reference
    g3 = mpp.Micropp3(mk(mpp, "homog_damage", 5, **kw))
    g3.read_restart(8)
    for gp in range(2):
        for o in (g, g2, g3, r):
            o.set_strain(gp, path[10][gp])
    for o in (g, g2, g3, r):
        o.homogenize()
    for gp in range(2):
        assert g2.is_non_linear(gp) == 1 and g3.is_non_linear(gp) == 1
        assert relerr(g2.get_stress(gp), g.get_stress(gp)) < 1e-12
        assert relerr(g3.get_stress(gp), r.get_stress(gp)) < 1e-7


def _vtu_arrays(path):
    """(structure, {name: values}) of an ASCII .vtu file: the text with every DataArray body blanked, and the numbers."""
    import re
    text = path.read_text()
    arrays = {}

    def grab(m):
        arrays[re.search(r'Name="(\w+)"', m.group(1)).group(1)] = np.array(m.group(2).split(), dtype=np.float64)
        return m.group(1) + "</DataArray>"

    skeleton = re.sub(r"(<DataArray[^>]*>)(.*?)</DataArray>", grab, text, flags=re.S)
    return skeleton, arrays


@pytest.mark.parametrize("case,n,steps", [("damage_sphere", 6, 10), ("plastic_layer", 5, 3), ("elastic_sphere", 5, 1)])
def test_vtu_output(mpp, refpy, tmp_path, case, n, steps):
    """output() writes the reference's VTU file (src/output.cpp:30-213): same XML skeleton byte for byte, same arrays
    (mesh/connectivity exactly; fields computed on the device by k_elem_fields within 1e-8)."""
    kw = dict(ngp=2, lin_stress=False, calc_ctan_lin=False, nr_max_its=10)
    g, r = pair(mpp, refpy, case, n, **kw)
    path = load_path(2, steps, 9)
    run_history(g, path)
    run_history(r, path)
    for gp in range(2):
        g.output(gp, tmp_path / f"ours{gp}")
        r.output(gp, tmp_path / f"ref{gp}")
        so, ao = _vtu_arrays(tmp_path / f"ours{gp}.vtu")
        sr, ar = _vtu_arrays(tmp_path / f"ref{gp}.vtu")
        assert so == sr
        assert set(ao) == set(ar) == {"Position", "connectivity", "offsets", "types", "displ", "strain", "stress",
                                      "elem_type", "plasticity", "damage_e", "damage_D", "hardening"}
        for name in ("Position", "connectivity", "offsets", "types", "elem_type"):
            assert np.array_equal(ao[name], ar[name]), name
        for name in ("displ", "strain", "stress", "plasticity", "damage_e", "damage_D", "hardening"):
            assert ao[name].shape == ar[name].shape, name
            assert relerr(ao[name], ar[name], floor=1e-30) < 2e-6, name  # the file holds 7 significant digits
    # output2 names the file itself (src/output.cpp:44-69)
    import os
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        g.output2(1, 42, 3)
    finally:
        os.chdir(cwd)
    assert _vtu_arrays(tmp_path / "micropp-42-3.vtu")[1]["displ"].shape == ao["displ"].shape


# ---------------------------------------------------------------- the bench workloads themselves vs the reference
def test_bench_elastic30_against_reference(mpp, refpy):
    """BASELINE configs[1] at FULL size (30^3, sphere, contrast 10, bench.py's seeded strains), three Gauss points
    through the default product path (implicit operator, TMA-tiled SpMV, CUDA graphs) against the compiled reference:
    stress to 1e-8 (north star), CG iterations +-1."""
    import bench
    wl = bench.WORKLOADS["elastic30"]
    ngp = 3
    eps = bench.strains_for("elastic30", 1024, 0, 0)[:ngp]
    g = mpp.Micropp3(mpp.default_params(size=(30, 30, 30), ngp=ngp, **wl["params"]))
    r = refpy.RefMicropp(refpy.default_params(size=(30, 30, 30), ngp=ngp, **wl["params"]))
    assert g.implicit_kernel() == 3          # k_spmv_dot_tmac (+ k_spmv_fix) is what the bench measures
    hg, hr = run_history(g, [eps]), run_history(r, [eps])
    compare_histories(hg, hr, newton_budget=1)
    assert 60 <= hg[0]["cost"][0] <= 90      # about 75 CG iterations at 30^3, contrast 10


def test_bench_damage_workload_against_reference(mpp, refpy):
    """BASELINE configs[2] (damage matrix + elastic sphere, nr_max_its = 12, bench.py's load path) at 20^3 -- the
    largest size the serial reference finishes in seconds -- over the load steps where the Gauss points turn
    non-linear: stress to 1e-8, same non-linear / converged flags, CG iterations +-1 per Newton step."""
    import bench
    wl = bench.WORKLOADS["damage50"]
    ngp = 2
    g = mpp.Micropp3(mpp.default_params(size=(20, 20, 20), ngp=ngp, **wl["params"]))
    r = refpy.RefMicropp(refpy.default_params(size=(20, 20, 20), ngp=ngp, **wl["params"]))
    path = [bench.strains_for("damage50", 512, 0, k)[[3, 200]] for k in range(7)]
    hg, hr = run_history(g, path), run_history(r, path)
    compare_histories(hg, hr, newton_budget=12)
    assert any(hg[-1]["nl"])                 # the path does reach the damage branch


@pytest.mark.parametrize("case,steps", [("damage_sphere", 7), ("elastic_sphere", 2), ("plastic_layer", 5),
                                        ("plastic_layer_yield", 7)])
def test_multi_wave_equals_single_wave(mpp, monkeypatch, case, steps):
    """More Gauss points than resident slots (MICROPP_WAVE caps the wave; at BASELINE sizes 4096 damage RVEs of 50^3 do
    not fit one GPU at once): the Gauss points are processed wave by wave with their FE state parked in HBM between
    waves -- results, costs and flags must be bit-identical to the all-resident run."""
    ngp = 7
    kw = dict(ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=12)
    path = load_path(ngp, steps, 99, comp=(1 if case == "plastic_layer" else 2 if case == "plastic_layer_yield" else 0),
                     eps_max=(1.0 if case == "plastic_layer" else 0.1))
    monkeypatch.setenv("MICROPP_WAVE", "3")
    a = mpp.Micropp3(mk(mpp, case, 7, **kw))
    assert a.wave_size() == 3
    monkeypatch.delenv("MICROPP_WAVE")
    b = mpp.Micropp3(mk(mpp, case, 7, **kw))
    assert b.wave_size() >= ngp
    ha, hb = run_history(a, path), run_history(b, path)
    for x, y in zip(ha, hb):
        assert np.array_equal(x["sig"], y["sig"])
        assert x["cost"] == y["cost"] and x["nl"] == y["nl"] and x["conv"] == y["conv"]
