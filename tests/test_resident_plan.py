"""CPU: the plan of the cluster-resident DPCG kernel (micropp_b200/csrc/cg_resident.cu) replayed on the host with the
kernel's own index functions -- per CTA a zeroed brick of p, own rows stored, halo rows delivered to the CTAs that
res_push_dests names, the chunk's pure-material row block on every node and the element-wise correction
(res_fix_corr) on the interface entries -- against an INDEPENDENT matrix-free FE product sum_e Ke_type(e) p_e written
in numpy.  Covers: every interior node produced exactly once (the replay returns -3 otherwise), the halo ring of every
CTA complete, the corner / element-position conventions of the correction, the fit of the plan into 227 KB.  The GPU
tests (tests/test_gpu_resident.py) then compare whole solves with the three-kernel loop and the reference."""
import ctypes as C

import numpy as np
import pytest

import micropp_b200 as M
from test_tiling import layer_types, sphere_types

CX = [0, 1, 1, 0, 0, 1, 1, 0]   # corner_x / corner_y / corner_z of fe_math.cuh (src/common.cpp:30-41)
CY = [0, 0, 1, 1, 0, 0, 1, 1]
CZ = [0, 0, 0, 0, 1, 1, 1, 1]


def corner_of(lx, ly, lz):
    return lz * 4 + ((3 - lx) if ly else lx)


def random_ke(rng):
    ke = np.zeros((3, 24, 24))
    for m in range(3):
        a = rng.standard_normal((24, 24))
        ke[m] = (a + a.T) * (1.0 + 4.0 * m) + 24 * (1 + m) * np.eye(24)
    return ke


def pure_rows(ke):
    """rows[m][nbr][fi*3+fj]: the gather of gather_block_elastic (mgpu_kernels.cu) for a node inside one material."""
    rows = np.zeros((3, 27, 9))
    for m in range(3):
        for nbr in range(27):
            di, dj, dk = nbr % 3 - 1, (nbr // 3) % 3 - 1, nbr // 9 - 1
            for c in range(8):
                ax, ay, az = (c >> 2) & 1, (c >> 1) & 1, c & 1
                lx, ly, lz = 1 - ax, 1 - ay, 1 - az
                mx, my, mz = lx + di, ly + dj, lz + dk
                if 0 <= mx <= 1 and 0 <= my <= 1 and 0 <= mz <= 1:
                    a, jn = corner_of(lx, ly, lz), corner_of(mx, my, mz)
                    rows[m, nbr] += ke[m, a * 3:a * 3 + 3, jn * 3:jn * 3 + 3].reshape(-1)
    return rows


def fe_product(dims, et, ke, p):
    """Ap = sum over elements of Ke p_e, interior rows only (p vanishes on the boundary).  p, Ap: [3][nz][ny][nx]."""
    nx, ny, nz = dims
    t = et.reshape(nz - 1, ny - 1, nx - 1)
    Ap = np.zeros_like(p)
    for a in range(8):
        for jn in range(8):
            blk = ke[t][..., a * 3:a * 3 + 3, jn * 3:jn * 3 + 3]                       # [ez][ey][ex][3][3]
            pj = p[:, CZ[jn]:CZ[jn] + nz - 1, CY[jn]:CY[jn] + ny - 1, CX[jn]:CX[jn] + nx - 1]
            contrib = np.einsum("zyxij,jzyx->izyx", blk, pj)
            Ap[:, CZ[a]:CZ[a] + nz - 1, CY[a]:CY[a] + ny - 1, CX[a]:CX[a] + nx - 1] += contrib
    Ap[:, 0], Ap[:, -1], Ap[:, :, 0], Ap[:, :, -1], Ap[:, :, :, 0], Ap[:, :, :, -1] = 0, 0, 0, 0, 0, 0
    return Ap


def replay(dims, et, ke, p, force_cs=0):
    lib = M.load()
    f = lib.mgpu_resident_replay_host
    f.restype = C.c_int
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    f.argtypes = [C.c_int] * 3 + [ip, dp, dp, dp, dp, ip, C.c_int]
    nx, ny, nz = dims
    et = np.ascontiguousarray(et, dtype=np.int32)
    rows = np.ascontiguousarray(pure_rows(ke))
    kef = np.ascontiguousarray(ke.reshape(-1))
    pf = np.ascontiguousarray(p.reshape(-1))
    Ap = np.zeros_like(pf)
    meta = np.zeros(8, dtype=np.int32)
    rc = f(nx, ny, nz, et.ctypes.data_as(ip), rows.ctypes.data_as(dp), kef.ctypes.data_as(dp), pf.ctypes.data_as(dp),
           Ap.ctypes.data_as(dp), meta.ctypes.data_as(ip), force_cs)
    return rc, Ap.reshape(p.shape), meta


CASES = [((30, 30, 30), "sphere", 0), ((20, 20, 20), "sphere", 0), ((10, 10, 10), "sphere", 0), ((5, 5, 5), "sphere", 0),
         ((3, 3, 3), "sphere", 0), ((11, 10, 13), "layer", 0), ((19, 12, 9), "layer", 0), ((24, 15, 9), "sphere", 0),
         ((16, 14, 21), "layer", 0), ((10, 6, 14), "sphere", 0), ((12, 12, 12), "layer", 2), ((12, 12, 12), "layer", 4),
         ((14, 12, 18), "sphere", 8), ((10, 10, 10), "sphere", 4), ((28, 28, 28), "sphere", 0), ((26, 30, 22), "layer", 0)]


@pytest.mark.parametrize("dims,kind,cs", CASES)
def test_replay_equals_matrix_free_product(dims, kind, cs):
    rng = np.random.default_rng(sum(dims) + cs)
    nx, ny, nz = dims
    et = sphere_types(*dims) if kind == "sphere" else layer_types(*dims)
    ke = random_ke(rng)
    p = rng.standard_normal((3, nz, ny, nx))
    p[:, 0], p[:, -1], p[:, :, 0], p[:, :, -1], p[:, :, :, 0], p[:, :, :, -1] = 0, 0, 0, 0, 0, 0
    rc, Ap, meta = replay(dims, et, ke, p, cs)
    if cs and rc == 0:
        pytest.skip(f"no plan with {cs} CTAs for {dims}")
    assert rc > 0, f"replay failed with {rc}"
    assert meta[0] == rc and meta[1] * meta[2] == rc and meta[5] <= 232448 and meta[4] % 32 == 0
    ref = fe_product(dims, et, ke, p)
    err = np.abs(Ap - ref).max() / np.abs(ref).max()
    assert err < 1e-13, err


def test_bench_size_plan():
    """30^3 (BASELINE configs[1]): 8 CTAs of 14 x 7 rows, 4 chunks of 7 nodes, 416 threads."""
    dims = (30, 30, 30)
    rng = np.random.default_rng(0)
    p = np.zeros((3, 30, 30, 30))
    rc, _, meta = replay(dims, sphere_types(*dims), random_ke(rng), p)
    assert rc == 8 and tuple(meta[:5]) == (8, 2, 4, 7, 416)


@pytest.mark.parametrize("n", [32, 50])
def test_too_large_for_a_cluster(n):
    """brick + du of a 32^3 RVE exceed 227 KB per CTA at 8 CTAs: such RVEs keep the three-kernel loop"""
    dims = (n, n, n)
    rc, _, _ = replay(dims, sphere_types(*dims), random_ke(np.random.default_rng(1)), np.zeros((3, n, n, n)))
    assert rc == 0


def hex8_ke(E, nu, dx, dy, dz):
    """24 x 24 stiffness of a hex8 element (2 x 2 x 2 Gauss points, isotropic material), node order of CX / CY / CZ."""
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    Cm = np.zeros((6, 6))
    Cm[:3, :3] = lam
    Cm[np.arange(3), np.arange(3)] += 2 * mu
    Cm[np.arange(3, 6), np.arange(3, 6)] = mu
    xi = np.array([[2 * CX[a] - 1, 2 * CY[a] - 1, 2 * CZ[a] - 1] for a in range(8)], dtype=float)
    ke = np.zeros((24, 24))
    g = 1 / np.sqrt(3.0)
    for gx in (-g, g):
        for gy in (-g, g):
            for gz in (-g, g):
                B = np.zeros((6, 24))
                for a in range(8):
                    dn = 0.125 * np.array([xi[a, 0] * (1 + xi[a, 1] * gy) * (1 + xi[a, 2] * gz) * 2 / dx,
                                           xi[a, 1] * (1 + xi[a, 0] * gx) * (1 + xi[a, 2] * gz) * 2 / dy,
                                           xi[a, 2] * (1 + xi[a, 0] * gx) * (1 + xi[a, 1] * gy) * 2 / dz])
                    B[0, 3 * a], B[1, 3 * a + 1], B[2, 3 * a + 2] = dn
                    B[3, 3 * a], B[3, 3 * a + 1] = dn[1], dn[0]
                    B[4, 3 * a], B[4, 3 * a + 2] = dn[2], dn[0]
                    B[5, 3 * a + 1], B[5, 3 * a + 2] = dn[2], dn[1]
                ke += B.T @ Cm @ B * (dx * dy * dz / 8)
    return ke


def test_mirror_symmetry_pattern_of_pure_row_blocks():
    """The 153-term operator copy of k_cg_resident rests on this: the ELL row block of a node surrounded by ONE isotropic
    material on a regular grid couples components fi != fj only towards neighbours whose offset is non-zero along both
    axes -- 90 of the 243 entries vanish (also for dx != dy != dz); a generic symmetric element matrix has no such
    pattern and gets the dense copy."""
    lib = M.load()
    f = lib.mgpu_resident_rows_sparse_host
    f.restype = C.c_int
    f.argtypes = [C.POINTER(C.c_double)]
    for dims in ((1.0, 1.0, 1.0), (0.3, 0.5, 0.2)):
        ke = np.stack([hex8_ke(1e7, 0.3, *dims), hex8_ke(1e8, 0.25, *dims), hex8_ke(3e6, 0.2, *dims)])
        rows = np.ascontiguousarray(pure_rows(ke))
        nz = 0
        for nbr in range(27):
            o = (nbr % 3 - 1, (nbr // 3) % 3 - 1, nbr // 9 - 1)
            for fi in range(3):
                for fj in range(3):
                    structural = fi == fj or (o[fi] != 0 and o[fj] != 0)
                    nz += structural
                    if not structural:
                        assert abs(rows[:, nbr, fi * 3 + fj]).max() <= 1e-13 * abs(rows).max(), (nbr, fi, fj)
        assert nz == 153
        assert f(rows.ctypes.data_as(C.POINTER(C.c_double))) == 1
    rows = np.ascontiguousarray(pure_rows(random_ke(np.random.default_rng(3))))
    assert f(rows.ctypes.data_as(C.POINTER(C.c_double))) == 0
