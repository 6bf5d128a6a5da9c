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
