"""-m gpu: the cluster-resident DPCG kernel (k_cg_resident, micropp_b200/csrc/cg_resident.cu: the whole solve of
src/ell.cpp:66-122 in one launch, one thread-block cluster per RVE, p / du in shared memory, r in registers, halo rows
and dot products through distributed shared memory) against
  (a) the three-kernel DPCG loop of the same library (MICROPP_RESIDENT=0): same operator, dot products summed in another
      fixed order => stresses agree to rounding amplified by the solve, iteration counts to +-1;
  (b) the reference CPU path: stress 1e-8, CG iterations +-1 (the north-star tolerances);
for every cluster size (1, 2, 4, 8 CTAs: halo pushes in y, in z and across corners), for RVEs with interface nodes in
one, two and three materials, over several waves of clusters, with and without CUDA graphs."""
import os

import numpy as np
import pytest

from common import CASES, EL, relerr

pytestmark = pytest.mark.gpu

ELASTIC = {
    "sphere": CASES["elastic_sphere"],
    "layers": dict(type=2, geo_params=(0.5, 0.0, 0.0, 0.0), materials=[EL(1e7), EL(6e7, 0.25), EL(1e7)]),
    "fibres3": dict(type=10, materials=[EL(1e7), EL(1e8), EL(4e6, 0.2)]),
    "homog": dict(type=0, materials=[EL(3e7, 0.25)] * 3),
}


def run(mod_cls, params, eps, env=None):
    env = env or {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        g = mod_cls(params)
        ngp = eps.shape[0]
        for gp in range(ngp):
            g.set_strain(gp, eps[gp])
        g.homogenize()
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    sig = np.array([g.get_stress(gp) for gp in range(ngp)])
    cost = [g.get_cost(gp) for gp in range(ngp)]
    conv = [g.has_converged(gp) for gp in range(ngp)]
    return g, sig, cost, conv


def params_of(mod, dims, ngp, kind, **kw):
    return mod.default_params(size=dims, ngp=ngp, lin_stress=False, calc_ctan_lin=False, **ELASTIC[kind], **kw)


SHAPES = [((12, 12, 12), "sphere", 19), ((9, 11, 10), "sphere", 8), ((14, 9, 8), "layers", 3), ((30, 7, 6), "sphere", 2),
          ((19, 12, 5), "layers", 5), ((3, 3, 3), "sphere", 1), ((4, 4, 4), "homog", 2), ((16, 16, 16), "fibres3", 9),
          ((20, 20, 20), "sphere", 4), ((24, 15, 9), "fibres3", 3), ((30, 30, 30), "sphere", 20),
          ((28, 30, 26), "layers", 3), ((11, 10, 13), "sphere", 40)]


@pytest.mark.parametrize("dims,kind,ngp", SHAPES)
def test_resident_equals_three_kernel_loop(mpp, dims, kind, ngp):
    rng = np.random.default_rng(101)
    eps = rng.uniform(-1e-3, 1e-3, (ngp, 6))
    gr, sr, cr, vr = run(mpp.Micropp3, params_of(mpp, dims, ngp, kind), eps)
    info = gr.resident_info()
    assert info is not None, "the cluster-resident kernel must serve this RVE"
    gl, sl, cl, vl = run(mpp.Micropp3, params_of(mpp, dims, ngp, kind), eps, {"MICROPP_RESIDENT": "0"})
    assert gl.resident_info() is None
    assert vr == vl and all(vr)
    assert all(abs(a - b) <= 1 for a, b in zip(cr, cl)), (cr, cl)
    tol = 1e-9 if len(set(dims)) == 1 else 1e-4   # anisotropic meshes: see test_implicit_tma_kernels_equal_explicit
    for gp in range(ngp):
        assert relerr(sr[gp], sl[gp]) < tol
    for gp in (0, ngp - 1):
        assert relerr(gr.get_u(gp), gl.get_u(gp)) < (1e-7 if len(set(dims)) == 1 else 1e-3)
    # the same strain on every slot => the same bits on every slot (fixed summation order inside and across the CTAs)
    eps1 = np.tile(eps[:1], (ngp, 1))
    _, s1, c1, _ = run(mpp.Micropp3, params_of(mpp, dims, ngp, kind), eps1)
    assert all(np.array_equal(s1[0], s1[gp]) for gp in range(ngp)) and len(set(c1)) == 1


@pytest.mark.parametrize("cs", [1, 2, 4, 8])
@pytest.mark.parametrize("dims,kind", [((14, 12, 18), "sphere"), ((12, 18, 12), "layers"), ((16, 16, 16), "fibres3")])
def test_every_cluster_size(mpp, dims, kind, cs):
    """MICROPP_RESIDENT_CS forces the number of CTAs per cluster: 1 (no halo), 2 and 4 (split in z, or in y and z), 8."""
    ngp = 5
    rng = np.random.default_rng(7 + cs)
    eps = rng.uniform(-1e-3, 1e-3, (ngp, 6))
    gr, sr, cr, vr = run(mpp.Micropp3, params_of(mpp, dims, ngp, kind), eps, {"MICROPP_RESIDENT_CS": str(cs)})
    info = gr.resident_info()
    if info is None:
        pytest.skip(f"no plan with {cs} CTAs for {dims}")
    assert info["cs"] == cs
    gl, sl, cl, vl = run(mpp.Micropp3, params_of(mpp, dims, ngp, kind), eps, {"MICROPP_RESIDENT": "0"})
    assert vr == vl and all(abs(a - b) <= 1 for a, b in zip(cr, cl))
    for gp in range(ngp):
        assert relerr(sr[gp], sl[gp]) < (1e-9 if len(set(dims)) == 1 else 1e-4)   # as above


@pytest.mark.parametrize("dims,kind,ngp", [((10, 10, 10), "sphere", 11), ((7, 8, 9), "layers", 4),
                                           ((16, 16, 16), "fibres3", 3), ((20, 20, 20), "sphere", 2)])
def test_resident_vs_reference(mpp, refpy, dims, kind, ngp):
    rng = np.random.default_rng(9)
    eps = rng.uniform(-1e-3, 1e-3, (ngp, 6))
    kw = dict(size=dims, ngp=ngp, lin_stress=False, calc_ctan_lin=True, **ELASTIC[kind])
    g, sg, cg, vg = run(mpp.Micropp3, mpp.default_params(**kw), eps)
    assert g.resident_info() is not None
    r, sr, cr, vr = run(refpy.RefMicropp, refpy.default_params(**kw), eps)
    assert vg == vr
    for gp in range(ngp):
        assert relerr(sg[gp], sr[gp]) < 1e-8          # north-star tolerance
        assert abs(cg[gp] - cr[gp]) <= 1              # CG iterations (one Newton step each)
    assert relerr(g.ctan_lin(), r.ctan_lin()) < 1e-8  # the 6 unit-strain solves of the constructor


def test_graphs_off_same_bits(mpp):
    dims, ngp = (14, 14, 14), 7
    rng = np.random.default_rng(3)
    eps = rng.uniform(-1e-3, 1e-3, (ngp, 6))
    _, s0, c0, _ = run(mpp.Micropp3, params_of(mpp, dims, ngp, "sphere"), eps)
    _, s1, c1, _ = run(mpp.Micropp3, params_of(mpp, dims, ngp, "sphere"), eps, {"MICROPP_GRAPHS": "0"})
    assert c0 == c1 and np.array_equal(s0, s1)


def test_second_load_step_and_zero_strain(mpp, refpy):
    """u_k of step 1 is the start of step 2 (few iterations); a zero strain needs no iteration at all: the kernel leaves
    du = 0 and cg_its = 0 (loop-head test of src/ell.cpp:93-94 on |z0| = 0 < cg_abs_tol)."""
    dims, ngp = (12, 12, 12), 4
    kw = dict(size=dims, ngp=ngp, lin_stress=False, calc_ctan_lin=False, **ELASTIC["sphere"])
    g = mpp.Micropp3(mpp.default_params(**kw))
    r = refpy.RefMicropp(refpy.default_params(**kw))
    rng = np.random.default_rng(5)
    for step in range(3):
        eps = rng.uniform(-1e-3, 1e-3, (ngp, 6)) * (step + 1)
        eps[1] = 0.0
        for m in (g, r):
            for gp in range(ngp):
                m.set_strain(gp, eps[gp])
            m.homogenize()
            m.update_vars()
        for gp in range(ngp):
            assert abs(g.get_cost(gp) - r.get_cost(gp)) <= 1
            assert relerr(g.get_stress(gp), r.get_stress(gp), floor=1e-3) < 1e-8
        assert g.get_cost(1) == 0 and np.all(np.asarray(g.get_stress(1)) == 0.0)


def test_residual_history_against_three_kernel_loop(mpp):
    """|z| at the head of every DPCG iteration, as recorded by both solvers (mgpu_cg_history)."""
    dims = (16, 16, 16)
    eps = np.array([[1e-3, -2e-4, 3e-4, 5e-4, 0.0, -1e-4]])
    hist = []
    for env in ({}, {"MICROPP_RESIDENT": "0"}):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            g = mpp.Micropp3(params_of(mpp, dims, 1, "sphere"))
        finally:
            for k, v in old.items():
                if v is None:
                    del os.environ[k]
        g.cg_history(64)
        g.set_strain(0, eps[0])
        g.homogenize()
        hist.append(g.cg_history_read(0, 64))
    a, b = hist
    n = min(len(a), len(b))
    assert abs(len(a) - len(b)) <= 1 and n > 10
    assert relerr(a[:12], b[:12]) < 1e-12
    assert np.all(np.abs(a[:n] - b[:n]) <= 1e-6 * np.abs(b[:n]) + 1e-9 * b[0])


@pytest.mark.parametrize("dims,kind", [((12, 12, 12), "sphere"), ((9, 11, 10), "fibres3"), ((14, 9, 8), "layers"),
                                       ((30, 30, 30), "sphere"), ((13, 7, 6), "sphere")])
def test_residual_through_the_operator(mpp, refpy, dims, kind):
    """assembly_rhs of an all-elastic RVE as b = -A u through the implicit operator (k_u_to_p, the implicit SpMV,
    k_rhs_from_ap) against the element loop of the same library (MICROPP_RHS_OPERATOR=0: k_elem_rhs + k_asm_rhs) and
    against the reference's assembly_rhs (src/assembly.cpp:61-104), for a random u WITH boundary values."""
    kw = dict(size=dims, ngp=1, lin_stress=False, calc_ctan_lin=False, **ELASTIC[kind])
    g = mpp.Micropp3(mpp.default_params(**kw))
    os.environ["MICROPP_RHS_OPERATOR"] = "0"
    try:
        ge = mpp.Micropp3(mpp.default_params(**kw))
    finally:
        del os.environ["MICROPP_RHS_OPERATOR"]
    r = refpy.RefMicropp(refpy.default_params(**kw))
    u = np.random.default_rng(11).uniform(-1e-3, 1e-3, g.nndim)
    bo, no = g.assembly_rhs(u)
    be, ne = ge.assembly_rhs(u)
    br, nr = r.assembly_rhs(u)
    assert relerr(bo, br) < 1e-12 and relerr(be, br) < 1e-12 and relerr(bo, be) < 1e-12
    assert abs(no - nr) <= 1e-12 * nr and abs(ne - nr) <= 1e-12 * nr


def test_two_live_contexts_share_a_kernel_instantiation(mpp):
    """Two live objects whose plans use the same kernel instantiation with different shared-memory sizes: the
    per-kernel dynamic shared memory attribute must only ever grow (the larger plan is solved after the smaller object
    was constructed)."""
    rng = np.random.default_rng(2)
    big = mpp.Micropp3(params_of(mpp, (14, 14, 14), 2, "sphere"))
    small = mpp.Micropp3(params_of(mpp, (12, 12, 12), 2, "sphere"))
    ib, is_ = big.resident_info(), small.resident_info()
    assert ib and is_ and ib["tn"] == is_["tn"] and (ib["threads"] <= 384) == (is_["threads"] <= 384)
    assert ib["smem"] > is_["smem"]
    for obj in (big, small, big):
        eps = rng.uniform(-1e-3, 1e-3, (2, 6))
        for gp in range(2):
            obj.set_strain(gp, eps[gp])
        obj.homogenize()
        assert all(obj.has_converged(gp) for gp in range(2))


@pytest.mark.parametrize("dims,kind", [((12, 12, 12), "sphere"), ((30, 30, 30), "sphere"), ((11, 10, 13), "fibres3")])
def test_dense_operator_copy(mpp, dims, kind):
    """MICROPP_RESIDENT_DENSE=1: the 243-term copy of the operator (taken when the pure-material row blocks do not have
    the mirror-symmetry zero pattern) against the default 153-term copy: the skipped terms are zero up to rounding."""
    ngp = 3
    eps = np.random.default_rng(17).uniform(-1e-3, 1e-3, (ngp, 6))
    _, s0, c0, v0 = run(mpp.Micropp3, params_of(mpp, dims, ngp, kind), eps)
    _, s1, c1, v1 = run(mpp.Micropp3, params_of(mpp, dims, ngp, kind), eps, {"MICROPP_RESIDENT_DENSE": "1"})
    assert v0 == v1 and all(v0) and all(abs(a - b) <= 1 for a, b in zip(c0, c1))
    for gp in range(ngp):
        assert relerr(s0[gp], s1[gp]) < (1e-9 if len(set(dims)) == 1 else 1e-4)
