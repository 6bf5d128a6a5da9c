"""CPU-side check of the arithmetic the CUDA kernels run: micropp_b200/csrc/fe_math.cuh compiled for the
host (test-only shared object) against the compiled reference (oracle/_ref).  Bit-exactness is expected
for the material laws, strains and boundary displacements (non-contracted arithmetic, same order)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "hostcheck" / "fe_math_host.cpp"
SO = ROOT / "tests" / "hostcheck" / "libfe_math_host.so"
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hc():
    hdr = ROOT / "micropp_b200" / "csrc" / "fe_math.cuh"
    if not SO.exists() or SO.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                               "-I", str(hdr.parent), str(SRC), "-o", str(SO)])
    return C.CDLL(str(SO))


def d(a):
    return a.ctypes.data_as(_dp)


MATS = [(0, 1e7, 0.3, 0, 0, 0), (1, 1e3, 0.3, 5e4, 1e3, 0), (1, 3e7, 0.25, 1e7, 1e5, 0), (2, 1e7, 0.3, 0, 0, 1e5),
        (2, 3e7, 0.25, 0, 0, 1e5)]


def margs(m):
    return [C.c_int(m[0])] + [C.c_double(float(v)) for v in m[1:]]


def rand_vars(rng, mtype):
    v = np.zeros(7)
    if mtype == 1:
        v[:6] = rng.uniform(-1e-3, 1e-3, 6)
        v[6] = rng.uniform(0, 1e-3)
    elif mtype == 2:
        v[0] = rng.uniform(10, 80)
        v[1] = rng.uniform(0, 0.7)
    return v


@pytest.mark.parametrize("m", MATS)
def test_material_laws_bit_exact(hc, refpy, m):
    rng = np.random.default_rng(m[0] * 7 + 1)
    for trial in range(200):
        eps = rng.uniform(-1, 1, 6) * 10.0 ** rng.uniform(-5, -1)
        for vars_old in (None, rand_vars(rng, m[0])):
            vp = None if vars_old is None else d(vars_old)
            sig, c = np.zeros(6), np.zeros(36)
            hc.hc_mat_stress(*margs(m), d(eps), vp, d(sig))
            assert np.array_equal(sig, refpy.mat_stress(m, eps, vars_old))
            hc.hc_mat_ctan(*margs(m), d(eps), vp, d(c))
            assert np.array_equal(c, refpy.mat_ctan(m, eps, vars_old))
            vn = np.zeros(7)
            nl = hc.hc_mat_evolute(*margs(m), d(eps), vp, d(vn))
            vr, nlr = refpy.mat_evolute(m, eps, vars_old)
            assert bool(nl) == nlr
            assert np.array_equal(vn, vr)


def test_strain_bit_exact(hc, refpy):
    r = refpy.RefMicropp(refpy.default_params(size=(4, 5, 6), calc_ctan_lin=False))
    B = r.bmat()
    rng = np.random.default_rng(0)
    for gp in range(8):
        dsh = np.zeros(24)
        for a in range(8):
            dsh[a * 3 + 0], dsh[a * 3 + 1], dsh[a * 3 + 2] = B[gp, 0, a * 3], B[gp, 1, a * 3 + 1], B[gp, 2, a * 3 + 2]
        # the B layout the kernels assume (src/micro3D.cpp:100-119)
        Bchk = np.zeros((6, 24))
        for a in range(8):
            gx, gy, gz = dsh[a * 3:a * 3 + 3]
            Bchk[0, a * 3], Bchk[1, a * 3 + 1], Bchk[2, a * 3 + 2] = gx, gy, gz
            Bchk[3, a * 3], Bchk[3, a * 3 + 1] = gy, gx
            Bchk[4, a * 3], Bchk[4, a * 3 + 2] = gz, gx
            Bchk[5, a * 3 + 1], Bchk[5, a * 3 + 2] = gz, gy
        assert np.array_equal(Bchk, B[gp])
        ue = rng.uniform(-1, 1, 24)
        eps = np.zeros(6)
        hc.hc_strain(d(dsh), d(ue), d(eps))
        ref = np.zeros(6)
        for v in range(6):  # src/common.cpp:67-71 order
            acc = 0.0
            for i in range(24):
                acc += B[gp, v, i] * ue[i]
            ref[v] = acc
        assert np.array_equal(eps, ref)


def test_boundary_displacement_bit_exact(hc, refpy):
    dims = (4, 6, 5)
    r = refpy.RefMicropp(refpy.default_params(size=dims, calc_ctan_lin=False))
    eps = np.array([1.1e-3, -2.3e-3, 3.7e-3, 1.3e-3, -0.7e-3, 0.9e-3])
    u = r.set_displ_bc(eps, np.full(r.nndim, 123.0)).reshape(-1, 3)
    nx, ny, nz = dims
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                n = (k * ny + j) * nx + i
                bnd = i in (0, nx - 1) or j in (0, ny - 1) or k in (0, nz - 1)
                if not bnd:
                    assert np.all(u[n] == 123.0)
                    continue
                u3 = np.zeros(3)
                hc.hc_bc(i, j, k, nx, ny, nz, d(eps), d(u3))
                assert np.array_equal(u3, u[n]), (i, j, k)


def test_scatter_map_equals_reference_table(hc, refpy):
    # cols_row[8][8] of src/ell-common.cpp:175-178, observed through the reference's own scatter
    nx = ny = nz = 3
    Ae = np.arange(576, dtype=np.float64).reshape(24, 24) + 1.0
    vals = refpy.ell_add_one(nx, ny, nz, 1, 1, 1, Ae)
    nodes = refpy.elem_nodes(nx, ny, 1, 1, 1)
    for a in range(8):
        for j in range(8):
            slot = hc.hc_cols_row(a, j)
            for fi in range(3):
                for fj in range(3):
                    assert vals[nodes[a] * 3 + fi, slot * 3 + fj] == Ae[a * 3 + fi, j * 3 + fj]
    # and nothing else was touched
    assert np.count_nonzero(vals) == 576
