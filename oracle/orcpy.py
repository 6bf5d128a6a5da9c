"""oracle/orcpy.py -- TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/liboracle.so`` (the plain-C restatement ``oracle/micropp_oracle.c`` of the reference's
RVE-homogenization hot path).  Only ``tests/``, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline leg may
import this module; the product package ``micropp_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
SRC = HERE / "micropp_oracle.c"

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Material(C.Structure):
    _fields_ = [("E", C.c_double), ("nu", C.c_double), ("Ka", C.c_double), ("Sy", C.c_double), ("k", C.c_double),
                ("mu", C.c_double), ("lam", C.c_double), ("Xt", C.c_double), ("type", C.c_int)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
            subprocess.check_call(["make", "-s", "-C", str(HERE), "restatement"])
        lib = C.CDLL(str(LIB))
        lib.orc_new.restype = C.c_void_p
        lib.orc_gp_new.restype = C.c_void_p
        lib.orc_assembly_rhs.restype = C.c_double
        _lib = lib
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _v(a):
    return None if a is None else _d(a)


def material(m):
    """m = (type, E, nu, Ka, Sy, Xt)"""
    s = Material()
    load().orc_material_set(C.byref(s), int(m[0]), *[C.c_double(float(v)) for v in m[1:6]])
    return s


def ell_cols(nx, ny, nz):
    cols = np.zeros((3 * nx * ny * nz, 81), dtype=np.int32)
    load().orc_ell_cols(nx, ny, nz, cols.ctypes.data_as(_ip))
    return cols


def elem_nodes(nx, ny, ex, ey, ez):
    n = np.zeros(8, dtype=np.int32)
    load().orc_elem_nodes(nx, ny, ex, ey, ez, n.ctypes.data_as(_ip))
    return n


def elem_colour(ex, ey, ez):
    return int(load().orc_elem_colour(ex, ey, ez))


def cols_row(i, j):
    return int(load().orc_cols_row(i, j))


def elem_types(micro_type, geo, dims):
    nx, ny, nz = dims
    g = (C.c_double * 4)(*[float(v) for v in geo])
    out = np.zeros((nx - 1) * (ny - 1) * (nz - 1), dtype=np.int32)
    lib = load()
    for ez in range(nz - 1):
        for ey in range(ny - 1):
            for ex in range(nx - 1):
                out[(ez * (ny - 1) + ey) * (nx - 1) + ex] = lib.orc_elem_type(int(micro_type), g, nx, ny, nz, ex, ey, ez)
    return out


def mvp3(m, x):
    y = np.zeros(3)
    load().orc_mvp3(_d(np.ascontiguousarray(m, dtype=np.float64)), _d(np.ascontiguousarray(x, dtype=np.float64)), _d(y))
    return y


def mat_stress(m, eps, vars_old=None):
    s = np.zeros(6)
    mm = material(m)
    load().orc_mat_stress(C.byref(mm), _d(np.ascontiguousarray(eps, dtype=np.float64)), _v(vars_old), _d(s))
    return s


def mat_ctan(m, eps, vars_old=None):
    c = np.zeros(36)
    mm = material(m)
    load().orc_mat_ctan(C.byref(mm), _d(np.ascontiguousarray(eps, dtype=np.float64)), _v(vars_old), _d(c))
    return c


def mat_evolute(m, eps, vars_old=None):
    vn = np.zeros(7)
    mm = material(m)
    nl = load().orc_mat_evolute(C.byref(mm), _d(np.ascontiguousarray(eps, dtype=np.float64)), _v(vars_old), _d(vn))
    return vn, bool(nl)


def ell_mvp(nx, ny, nz, vals, x):
    y = np.zeros(3 * nx * ny * nz)
    load().orc_ell_mvp(nx, ny, nz, _d(np.ascontiguousarray(vals)), _d(np.ascontiguousarray(x)), _d(y))
    return y


def ell_solve_cgpd(nx, ny, nz, vals, b):
    x = np.zeros(3 * nx * ny * nz)
    err = C.c_double(0.0)
    its = load().orc_ell_solve_cgpd(nx, ny, nz, _d(np.ascontiguousarray(vals)), _d(np.ascontiguousarray(b)), _d(x),
                                    C.byref(err))
    return x, int(its), float(err.value)


def ell_solve_cgpd_hist(nx, ny, nz, vals, b, k):
    """ell_solve_cgpd + hist[i] = |z| seen by the loop-head test of iteration i (i < k)."""
    x = np.zeros(3 * nx * ny * nz)
    err = C.c_double(0.0)
    hist = np.zeros(k)
    its = load().orc_ell_solve_cgpd_hist(nx, ny, nz, _d(np.ascontiguousarray(vals)), _d(np.ascontiguousarray(b)), _d(x),
                                         C.byref(err), _d(hist), int(k))
    return x, int(its), hist[:min(k, int(its) + 1)]


class OrcProblem:
    """One RVE mesh + materials (the FE stages of the reference's micropp<3>)."""

    def __init__(self, params: dict, elem_type=None):
        self.lib = load()
        self.p = params
        self.nx, self.ny, self.nz = [int(v) for v in params["size"]]
        self.nn = self.nx * self.ny * self.nz
        self.nndim = 3 * self.nn
        self.nelem = (self.nx - 1) * (self.ny - 1) * (self.nz - 1)
        self.nvars = self.nelem * 56
        if elem_type is None:
            elem_type = elem_types(params["type"], params["geo_params"], params["size"])
            if np.any(elem_type < 0):
                raise ValueError("micro-structure not restated in the oracle: pass elem_type explicitly")
        self._et = np.ascontiguousarray(elem_type, dtype=np.int32)
        mats = (Material * 3)(*[material(m) for m in params["materials"][:3]])
        self._mats = mats
        self.h = C.c_void_p(self.lib.orc_new(self.nx, self.ny, self.nz, self._et.ctypes.data_as(_ip), mats,
                                             int(params.get("nr_max_its", 4)),
                                             C.c_double(float(params.get("nr_max_tol", 1e-10))),
                                             C.c_double(float(params.get("nr_rel_tol", 1e-3)))))
        self.lin_stress = bool(params.get("lin_stress", True))

    def elem_type(self):
        return self._et.copy()

    def bmat(self):
        out = np.zeros((8, 6, 24))
        self.lib.orc_get_bmat(self.h, _d(out))
        return out

    def set_displ_bc(self, eps, u=None):
        u = np.zeros(self.nndim) if u is None else np.ascontiguousarray(u, dtype=np.float64).copy()
        self.lib.orc_set_displ_bc(self.h, _d(np.ascontiguousarray(eps, dtype=np.float64)), _d(u))
        return u

    def assembly_rhs(self, u, vars_old=None):
        b = np.zeros(self.nndim)
        nrm = self.lib.orc_assembly_rhs(self.h, _d(np.ascontiguousarray(u)), _v(vars_old), _d(b))
        return b, float(nrm)

    def assembly_mat(self, u, vars_old=None):
        vals = np.zeros((self.nndim, 81))
        self.lib.orc_assembly_mat(self.h, _d(np.ascontiguousarray(u)), _v(vars_old), _d(vals))
        return vals

    def newton(self, eps, u, vars_old=None):
        u = np.ascontiguousarray(u, dtype=np.float64).copy()
        out = np.zeros(3, dtype=np.int32)
        self.lib.orc_newton(self.h, _d(np.ascontiguousarray(eps, dtype=np.float64)), _v(vars_old), _d(u),
                            out.ctypes.data_as(_ip))
        return u, dict(its=int(out[0]), solver_its=int(out[1]), converged=bool(out[2]))

    def ave_stress(self, u, vars_old=None):
        s = np.zeros(6)
        self.lib.orc_ave_stress(self.h, _d(np.ascontiguousarray(u)), _v(vars_old), _d(s))
        return s

    def vars_new(self, u, vars_old=None):
        vn = np.zeros(self.nvars)
        nl = self.lib.orc_vars_new(self.h, _d(np.ascontiguousarray(u)), _v(vars_old), _d(vn))
        return vn, bool(nl)


class OrcMicropp(OrcProblem):
    """ngp Gauss points, FE_ONE_WAY without sub-iterations: the public API subset the parity tests drive."""

    def __init__(self, params: dict, elem_type=None, ctan_lin=None):
        super().__init__(params, elem_type)
        self.ngp = int(params.get("ngp", 1))
        self.gps = [C.c_void_p(self.lib.orc_gp_new(self.h)) for _ in range(self.ngp)]
        if ctan_lin is not None:
            c = np.ascontiguousarray(ctan_lin, dtype=np.float64)
            for g in self.gps:
                self.lib.orc_gp_set_ctan(g, _d(c))

    def set_strain(self, gp, eps):
        self.lib.orc_gp_set_strain(self.gps[gp], _d(np.ascontiguousarray(eps, dtype=np.float64)))

    def get_stress(self, gp):
        s = np.zeros(6)
        self.lib.orc_gp_get_stress(self.gps[gp], _d(s))
        return s

    def homogenize(self):
        for g in self.gps:
            self.lib.orc_gp_homogenize(self.h, g, int(self.lin_stress))

    def update_vars(self):
        for g in self.gps:
            self.lib.orc_gp_update_vars(g)

    def get_cost(self, gp):
        return int(self.lib.orc_gp_cost(self.gps[gp]))

    def has_converged(self, gp):
        return bool(self.lib.orc_gp_converged(self.gps[gp]))

    def is_non_linear(self, gp):
        return int(self.lib.orc_gp_allocated(self.gps[gp]))
