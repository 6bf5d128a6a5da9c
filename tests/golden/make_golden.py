"""Generate the golden fixtures of tests/golden/ from the UNMODIFIED reference CPU implementation.

Run in the build container (needs /root/reference; `make -C oracle ref` first):

    python tests/golden/make_golden.py

The compiled reference (oracle/_ref/libmicropp_ref.so = reference sources + oracle/ref_shim.cpp) is driven through
oracle/refpy.py on small seeded inputs; inputs and outputs are stored as compressed .npz files small enough to commit.
The fixtures pin what the reference's own tests leave unpinned: the 3-D ELL column table, element
classification, every FE stage on a heterogeneous RVE, CG / Newton iteration counts and whole homogenize() histories.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from common import CASES, random_u, random_vars  # noqa: E402
from oracle import refpy as R  # noqa: E402


def mat_types(ref):
    et = ref.elem_type()
    mtypes = [m[0] for m in ref.p["materials"]]
    return [mtypes[t] for t in et]


def stage_fixture(case, dims, seed):
    p = R.default_params(size=dims, calc_ctan_lin=False, **CASES[case])
    r = R.RefMicropp(p)
    u = random_u(r.nndim, seed, 5e-3)
    v = random_vars(r.nelem, mat_types(r), seed + 1)
    eps = np.array([1.1e-3, -2.3e-3, 3.7e-3, 1.3e-3, -0.7e-3, 0.9e-3])
    out = dict(dims=np.array(dims), elem_type=r.elem_type(), bmat=r.bmat(), u=u, vars=v, eps=eps,
               u_bc=r.set_displ_bc(eps, u))
    for tag, vv in (("nov", None), ("v", v)):
        b, nrm = r.assembly_rhs(u, vv)
        out[f"b_{tag}"], out[f"bnorm_{tag}"] = b, nrm
        out[f"A_{tag}"] = r.assembly_mat(u, vv)
        out[f"sig_{tag}"] = r.ave_stress(u, vv)
        vn, nl = r.vars_new(u, vv)
        out[f"vnew_{tag}"], out[f"nl_{tag}"] = vn, nl
    # one DPCG solve on the assembled system (the structure of test/test_cg.cpp:46-80)
    e2 = np.array([1.0, 2.0, 3.0, 1.0, 1.0, 1.0]) * 1e-3
    u2 = r.set_displ_bc(e2)
    A2 = r.assembly_mat(u2)
    b2, _ = r.assembly_rhs(u2)
    x, its, err = R.ell_solve_cgpd(*dims, A2, b2)
    xs = random_u(r.nndim, seed + 2, 1.0)
    out.update(cg_A=A2, cg_b=b2, cg_x=x, cg_its=its, cg_err=err, mvp_x=xs, mvp_y=R.ell_mvp(*dims, A2, xs))
    # a Newton solve from u = 0
    e3 = np.array([0.01, -0.004, 0.002, 0.006, -0.003, 0.001]) * (0.1 if case == "elastic_sphere" else 1.0)
    un, st = r.newton(e3, np.zeros(r.nndim))
    out.update(nr_eps=e3, nr_u=un, nr_its=st["its"], nr_solver_its=st["solver_its"], nr_conv=st["converged"],
               nr_sig=r.ave_stress(un))
    r.close()
    if case in MUST_BE_NON_LINEAR:
        assert out["nl_v"] and out["nl_nov"], f"{case}: the stage fixture must exercise the non-linear material branch"
    return out


MUST_BE_NON_LINEAR = ("damage_sphere", "plastic_layer_yield", "mic3d_8")


def history_fixture(case, n, ngp, steps, seed, comp, nr_max_its, eps_max=0.1, dt=0.015, reverse_after=None, inc=None):
    p = R.default_params(size=(n, n, n), ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=nr_max_its,
                         **CASES[case])
    r = R.RefMicropp(p)
    rng = np.random.default_rng(seed)
    scale = rng.uniform(0.5, 1.5, ngp)
    eps_hist, sig, cost, conv, nl = [], [], [], [], []
    e = np.zeros((ngp, 6))
    for k in range(steps):
        if inc is None:
            e = np.zeros((ngp, 6))
            e[:, comp] = scale * eps_max * dt * k
        else:
            e = e.copy()
            e[:, comp] += (inc if (reverse_after is None or k < reverse_after) else -inc) * scale
        for g in range(ngp):
            r.set_strain(g, e[g])
        r.homogenize()
        eps_hist.append(e.copy())
        sig.append([r.get_stress(g) for g in range(ngp)])
        cost.append([r.get_cost(g) for g in range(ngp)])
        conv.append([r.has_converged(g) for g in range(ngp)])
        nl.append([r.is_non_linear(g) for g in range(ngp)])
        r.update_vars()
    out = dict(n=n, ngp=ngp, nr_max_its=nr_max_its, eps=np.array(eps_hist), sig=np.array(sig), cost=np.array(cost),
               conv=np.array(conv), nl=np.array(nl), elem_type=r.elem_type())
    r.close()
    if case in MUST_BE_NON_LINEAR:
        assert out["nl"][-1].all(), f"{case}: every Gauss point of the history must end non-linear"
    return out


def main():
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    np.savez_compressed(HERE / "ell_cols_3x4x5.npz", cols=R.ell_cols(3, 4, 5))
    np.savez_compressed(HERE / "ell_cols_2x2x2.npz", cols=R.ell_cols(2, 2, 2))
    # element classification of all 13 micro-structures on an anisotropic grid
    et = {}
    for mt in range(13):
        p = R.default_params(size=(9, 11, 10), type=mt, geo_params=(0.2, 0.1, 0.1, 0.1), calc_ctan_lin=False)
        r = R.RefMicropp(p)
        et[f"type{mt}"] = r.elem_type()
        r.close()
    np.savez_compressed(HERE / "elem_type_9x11x10.npz", **et)
    for case, dims, seed in (("damage_sphere", (5, 6, 4), 11), ("plastic_layer", (4, 5, 6), 21),
                             ("elastic_sphere", (6, 5, 5), 31), ("mic3d_8", (6, 6, 5), 41),
                             ("plastic_layer_yield", (5, 6, 4), 51)):
        np.savez_compressed(HERE / f"stages_{case}.npz", **stage_fixture(case, dims, seed))
    np.savez_compressed(HERE / "history_damage_sphere.npz",
                        **history_fixture("damage_sphere", 6, 3, 8, 1234, 0, 12))
    np.savez_compressed(HERE / "history_plastic_layer.npz",
                        **history_fixture("plastic_layer", 6, 2, 10, 7, 1, 8, inc=0.01, reverse_after=6))
    # J2 plasticity that really yields: load for 7 steps, then unload (test/test3d_4.cpp:80-85 pattern) in eps_33
    np.savez_compressed(HERE / "history_plastic_layer_yield.npz",
                        **history_fixture("plastic_layer_yield", 7, 3, 12, 17, 2, 12, inc=0.0015, reverse_after=7))
    np.savez_compressed(HERE / "history_elastic_sphere.npz",
                        **history_fixture("elastic_sphere", 8, 4, 2, 5, 0, 4, eps_max=0.01))
    # linear homogenized tangent of the constructor (src/micropp.cpp:256-284)
    p = R.default_params(size=(6, 6, 6), **CASES["damage_sphere"])
    r = R.RefMicropp(p)
    np.savez_compressed(HERE / "ctan_lin_damage_sphere_6.npz", ctan_lin=r.ctan_lin())
    r.close()
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
