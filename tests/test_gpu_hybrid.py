"""-m gpu: the HYBRID operator of RVEs with a damage / plastic phase (mgpu_kernels.cu: k_probe_lin, k_hyb_list,
k_spmv_hyb) -- implicit elastic row blocks for every node, explicit ELL rows only for the nodes that touch an element
past its linear regime.  It replaces the fully assembled Jacobian of src/assembly.cpp:189-257 + the SpMV of
src/ell.cpp:38-57 wherever it applies (even nx, not the z-slab mode), so every case here runs against the compiled
reference AND against the product's own fully assembled path (MICROPP_HYBRID=0).
"""
import numpy as np
import pytest

from common import CASES, relerr
from test_gpu_parity import compare_histories, mk, run_history

pytestmark = pytest.mark.gpu


def _path(case, ngp, steps, comp):
    scale = np.random.default_rng(23).uniform(0.5, 1.5, ngp)
    inc = 0.004 if case.startswith("damage") or case == "mic3d_8" else 0.0015
    path, e = [], np.zeros((ngp, 6))
    for k in range(steps):
        e = e.copy()
        e[:, comp] += (inc if k < steps - 3 else -inc) * scale   # load, then three unloading steps
        path.append(e)
    return path


def _run_prof(m, path):
    m.prof_enable(True)
    m.prof_read(True)
    h = run_history(m, path)
    pr = m.prof_read(True)
    m.prof_enable(False)
    return h, pr


@pytest.mark.parametrize("case,comp", [("damage_sphere", 0), ("damage_sphere", 3), ("plastic_layer_yield", 2),
                                       ("plastic_fibre_yield", 1), ("mic3d_8", 0)])
def test_hybrid_operator_vs_reference_and_assembled(mpp, refpy, monkeypatch, case, comp):
    ngp, n = 4, 8
    kw = dict(ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=12)
    path = _path(case, ngp, 9, comp)
    hr = run_history(refpy.RefMicropp(mk(refpy, case, n, **kw)), path)
    assert any(hr[-1]["nl"])

    g = mpp.Micropp3(mk(mpp, case, n, **kw))
    assert g.hybrid_available()
    hg, pr = _run_prof(g, path)
    g.close()
    # the hybrid operator really served DPCG iterations, with a proper sub-set of the rows
    assert pr["hybrid_slot_apps"] > 0
    assert 0 < pr["hybrid_row_apps"] < pr["hybrid_slot_apps"] * (n - 2) ** 3
    compare_histories(hg, hr, newton_budget=12)

    monkeypatch.setenv("MICROPP_HYBRID", "0")
    a = mpp.Micropp3(mk(mpp, case, n, **kw))
    assert not a.hybrid_available()
    ha, pa = _run_prof(a, path)
    a.close()
    assert pa["hybrid_slot_apps"] == 0 and pa["spmv_slot_apps"] > 0
    compare_histories(ha, hr, newton_budget=12)
    for x, y in zip(hg, ha):
        # same rows, same value order per row; only the implicit rows differ in how their 81 products are summed
        assert x["cost"] == y["cost"] and x["nl"] == y["nl"] and x["conv"] == y["conv"]
        assert relerr(x["sig"], y["sig"]) < 1e-10


def test_hybrid_threshold_extremes(mpp, monkeypatch):
    """MICROPP_HYBRID_MAX = 0: no slot may list a row => every non-linear solve is fully assembled; = 1: every slot
    stays hybrid whatever its list.  Same results."""
    ngp, n, case = 3, 8, "damage_sphere"
    kw = dict(ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=12)
    path = _path(case, ngp, 8, 0)
    out = {}
    for frac in ("0", "1"):
        monkeypatch.setenv("MICROPP_HYBRID_MAX", frac)
        m = mpp.Micropp3(mk(mpp, case, n, **kw))
        out[frac] = _run_prof(m, path)
        m.close()
    (h0, p0), (h1, p1) = out["0"], out["1"]
    assert p1["hybrid_slot_apps"] > p0["hybrid_slot_apps"]
    assert p1["spmv_slot_apps"] < p0["spmv_slot_apps"]
    for x, y in zip(h0, h1):
        assert x["nl"] == y["nl"] and x["conv"] == y["conv"]
        assert relerr(x["sig"], y["sig"]) < 1e-9


def test_hybrid_graphs_off_is_bit_identical(mpp, monkeypatch):
    ngp, n, case = 3, 8, "plastic_layer_yield"
    kw = dict(ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=12)
    path = _path(case, ngp, 7, 2)
    m = mpp.Micropp3(mk(mpp, case, n, **kw))
    h1 = run_history(m, path)
    m.close()
    monkeypatch.setenv("MICROPP_GRAPHS", "0")
    m = mpp.Micropp3(mk(mpp, case, n, **kw))
    h0 = run_history(m, path)
    m.close()
    for x, y in zip(h0, h1):
        assert x["cost"] == y["cost"]
        assert np.array_equal(x["sig"], y["sig"])


def test_hybrid_not_offered_for_odd_nx_or_all_elastic(mpp):
    # the TMA tile kernel of the implicit operator needs an even nx (16-byte aligned row starts)
    kw = dict(ngp=1, lin_stress=False, calc_ctan_lin=False)
    m = mpp.Micropp3(mk(mpp, "damage_sphere", 7, **kw))
    assert not m.hybrid_available()
    m.close()
    m = mpp.Micropp3(mk(mpp, "elastic_sphere", 8, **kw))
    assert not m.hybrid_available()     # nothing to list: the implicit operator alone
    m.close()


def test_hybrid_fe_full_tangent_and_mixed_coupling(mpp, refpy):
    """FE_FULL perturbation solves (src/homogenize.cpp:176-210, 252-276) go through the same operator choice; settings
    and tolerance of test_gpu_parity.py::test_fe_full_ctan_values (both sides solve to the limit)."""
    ngp, n = 3, 8
    cpl = [mpp.FE_FULL, mpp.FE_ONE_WAY, mpp.FE_FULL]
    kw = dict(ngp=ngp, coupling=cpl, lin_stress=False, calc_ctan_lin=True, nr_max_its=20, nr_rel_tol=1e-10)
    for case, comp in (("damage_sphere", 0), ("plastic_layer_yield", 2)):
        path = _path(case, ngp, 6, comp)[:4]      # loading only
        g = mpp.Micropp3(mk(mpp, case, n, **kw))
        r = refpy.RefMicropp(mk(refpy, case, n, **kw))
        hg, pr = _run_prof(g, path)
        hr = run_history(r, path)
        assert pr["hybrid_slot_apps"] > 0 and hr[-1]["nl"][0]
        worst = 0.0
        for k, (a, b) in enumerate(zip(hg, hr)):
            assert a["nl"] == b["nl"] and a["conv"] == b["conv"], k
            for gp in range(ngp):
                assert relerr(a["sig"][gp], b["sig"][gp]) < 1e-10, (case, k, gp)
                worst = max(worst, relerr(a["ctan"][gp], b["ctan"][gp]))
        print(f"hybrid FE_FULL ctan {case}: worst relative difference to the reference {worst:.3e}")
        assert worst < 1e-8, (case, worst)
        g.close()


def test_hybrid_subiterations_with_default_linear_shortcuts(mpp, refpy):
    """Sub-iterations (src/homogenize.cpp:137-166: the strain increment applied in nsubiterations pieces when Newton
    fails) with a small Newton budget and the constructor's defaults (lin_stress, calc_ctan_lin = true) -- the hybrid
    operator is chosen anew in every Newton solve of every sub-step."""
    ngp, n = 3, 8
    kw = dict(ngp=ngp, coupling=[mpp.FE_ONE_WAY] * ngp, nr_max_its=2, subiterations=True, nsubiterations=3)
    path = _path("damage_sphere", ngp, 8, 0)
    g = mpp.Micropp3(mk(mpp, "damage_sphere", n, **kw))
    r = refpy.RefMicropp(mk(refpy, "damage_sphere", n, **kw))
    assert g.hybrid_available()
    hg, pr = _run_prof(g, path)
    hr = run_history(r, path)
    assert pr["hybrid_slot_apps"] > 0 and any(any(h["sub"]) for h in hr)   # sub-iterations really happened
    compare_histories(hg, hr, newton_budget=40)


def test_hybrid_multi_wave_and_restart(mpp, monkeypatch, tmp_path):
    """Waves smaller than ngp (FE state parked between waves) and a restart file in the middle of the load path: both
    must continue bit-identically -- the lists of the hybrid operator are derived state, rebuilt from u and the internal
    variables at every Newton iteration."""
    monkeypatch.chdir(tmp_path)
    ngp, n, case = 5, 8, "damage_sphere"
    kw = dict(ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=12)
    path = _path(case, ngp, 8, 0)
    monkeypatch.setenv("MICROPP_WAVE", "2")
    a = mpp.Micropp3(mk(mpp, case, n, **kw))
    assert a.wave_size() == 2
    monkeypatch.delenv("MICROPP_WAVE")
    b = mpp.Micropp3(mk(mpp, case, n, **kw))
    ha, hb = run_history(a, path[:5]), run_history(b, path[:5])
    assert any(hb[-1]["nl"])
    b.write_restart(3)
    c = mpp.Micropp3(mk(mpp, case, n, **kw))
    c.read_restart(3)
    ha += run_history(a, path[5:])
    hb += run_history(b, path[5:])
    hc = run_history(c, path[5:])
    for x, y in zip(ha, hb):
        assert np.array_equal(x["sig"], y["sig"])
        assert x["cost"] == y["cost"] and x["nl"] == y["nl"] and x["conv"] == y["conv"]
    for x, y in zip(hc, hb[5:]):
        # Gauss points that were still linear when the file was written restart from u = 0 instead of their
        # converged u_n (the reference's format stores u only with the internal variables): same solution, other path
        for gp in range(ngp):
            if hb[4]["nl"][gp]:
                assert np.array_equal(x["sig"][gp], y["sig"][gp]) and x["cost"][gp] == y["cost"][gp]
            else:
                assert relerr(x["sig"][gp], y["sig"][gp]) < 1e-4      # two DPCG solves to a 1e-5 reduction
        assert x["nl"] == y["nl"]
