"""oracle/refpy.py -- TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/_ref/libmicropp_ref*.so``: the UNMODIFIED reference CPU
implementation compiled from /root/reference by ``oracle/Makefile`` plus the C-ABI
window ``oracle/ref_shim.cpp``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module;
the product package ``micropp_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SERIAL = HERE / "_ref" / "libmicropp_ref.so"
REF_OMP = HERE / "_ref" / "libmicropp_ref_omp.so"

# enums of include/types.hpp:85-117 and include/material_base.h:27
MIC = dict(HOMOGENEOUS=0, SPHERE=1, LAYER_Y=2, CILI_FIB_X=3, CILI_FIB_Z=4, CILI_FIB_XZ=5, QUAD_FIB_XYZ=6,
           QUAD_FIB_XZ=7, QUAD_FIB_XZ_BROKEN_X=8, SPHERES=9, MIC3D_8=10, FIBS_20_ORDER=11, FIBS_20_DISORDER=12)
FE_LINEAR, FE_ONE_WAY, FE_FULL, MIX_RULE_CHAMIS = 0, 1, 2, 3
ELASTIC, PLASTIC, DAMAGE = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class RefParams(C.Structure):
    _fields_ = [
        ("ngp", C.c_int), ("size", C.c_int * 3), ("type", C.c_int), ("geo_params", C.c_double * 4),
        ("mat_type", C.c_int * 3), ("mat_E", C.c_double * 3), ("mat_nu", C.c_double * 3),
        ("mat_Ka", C.c_double * 3), ("mat_Sy", C.c_double * 3), ("mat_Xt", C.c_double * 3),
        ("coupling", _ip), ("subiterations", C.c_int), ("nsubiterations", C.c_int), ("mpi_rank", C.c_int),
        ("nr_max_its", C.c_int), ("nr_max_tol", C.c_double), ("nr_rel_tol", C.c_double),
        ("calc_ctan_lin", C.c_int), ("use_A0", C.c_int), ("its_with_A0", C.c_int), ("lin_stress", C.c_int),
        ("write_log", C.c_int),
    ]


def default_params(**kw) -> dict:
    """Defaults of micropp_params_t (include/types.hpp:43-83)."""
    p = dict(ngp=1, size=(5, 5, 5), type=0, geo_params=(0.1, 0.1, 0.1, 0.1),
             materials=[(0, 1.0e7, 0.3, 0.0, 0.0, 0.0)] * 3,  # (type, E, nu, Ka, Sy, Xt)
             coupling=None, subiterations=False, nsubiterations=10, mpi_rank=0, nr_max_its=4, nr_max_tol=1.0e-10,
             nr_rel_tol=1.0e-3, calc_ctan_lin=True, use_A0=False, its_with_A0=1, lin_stress=True, write_log=False)
    p.update(kw)
    return p


def fill_struct(S, p: dict, keep: list):
    s = S()
    s.ngp = int(p["ngp"])
    s.size[:] = [int(v) for v in p["size"]]
    s.type = int(p["type"])
    s.geo_params[:] = [float(v) for v in p["geo_params"]]
    for i, m in enumerate(p["materials"][:3]):
        s.mat_type[i] = int(m[0])
        s.mat_E[i], s.mat_nu[i], s.mat_Ka[i], s.mat_Sy[i], s.mat_Xt[i] = [float(v) for v in m[1:6]]
    if p.get("coupling") is not None:
        arr = np.ascontiguousarray(p["coupling"], dtype=np.int32)
        keep.append(arr)
        s.coupling = arr.ctypes.data_as(_ip)
    for k in ("subiterations", "nsubiterations", "mpi_rank", "nr_max_its", "calc_ctan_lin", "use_A0",
              "its_with_A0", "lin_stress", "write_log"):
        setattr(s, k, int(p[k]))
    s.nr_max_tol = float(p["nr_max_tol"])
    s.nr_rel_tol = float(p["nr_rel_tol"])
    return s


def _d(a):
    return a.ctypes.data_as(_dp)


def _vars_ptr(v):
    return None if v is None else _d(v)


_libs = {}


def load(omp: bool = False):
    path = REF_OMP if omp else REF_SERIAL
    if not path.exists():
        raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
    key = str(path)
    if key not in _libs:
        lib = C.CDLL(key, mode=os.RTLD_LOCAL)  # RTLD_LOCAL: keep reference symbols away from the product's
        lib.ref_new.restype = C.c_void_p
        lib.ref_new.argtypes = [C.POINTER(RefParams)]
        lib.ref_assembly_rhs.restype = C.c_double
        lib.ref_wg.restype = C.c_double
        for name in ("ref_free", "ref_homogenize", "ref_homogenize_linear", "ref_update_vars"):
            getattr(lib, name).argtypes = [C.c_void_p]
            getattr(lib, name).restype = None
        _libs[key] = lib
    return _libs[key]


def available(omp: bool = False) -> bool:
    return (REF_OMP if omp else REF_SERIAL).exists()


class RefMicropp:
    """The reference ``micropp<3>`` (include/micropp.hpp:54-218) behind the shim."""

    def __init__(self, params: dict, omp: bool = False):
        self.lib = load(omp)
        self._keep = []
        self.p = params
        rp = fill_struct(RefParams, params, self._keep)
        self.h = C.c_void_p(self.lib.ref_new(C.byref(rp)))
        self.ngp = params["ngp"]
        self.nx, self.ny, self.nz = params["size"]
        self.nn = self.nx * self.ny * self.nz
        self.nndim = 3 * self.nn
        self.nelem = (self.nx - 1) * (self.ny - 1) * (self.nz - 1)
        self.nvars = self.nelem * 56

    def close(self):
        if self.h:
            self.lib.ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # public API
    def set_strain(self, gp, eps):
        e = np.ascontiguousarray(eps, dtype=np.float64)
        self.lib.ref_set_strain(self.h, int(gp), _d(e))

    def get_stress(self, gp):
        out = np.zeros(6)
        self.lib.ref_get_stress(self.h, int(gp), _d(out))
        return out

    def get_ctan(self, gp):
        out = np.zeros(36)
        self.lib.ref_get_ctan(self.h, int(gp), _d(out))
        return out

    def homogenize(self):
        self.lib.ref_homogenize(self.h)

    def homogenize_linear(self):
        self.lib.ref_homogenize_linear(self.h)

    def update_vars(self):
        self.lib.ref_update_vars(self.h)

    def is_non_linear(self, gp):
        return int(self.lib.ref_is_non_linear(self.h, int(gp)))

    def get_non_linear_gps(self):
        return int(self.lib.ref_get_non_linear_gps(self.h))

    def get_cost(self, gp):
        return int(self.lib.ref_get_cost(self.h, int(gp)))

    def has_converged(self, gp):
        return bool(self.lib.ref_has_converged(self.h, int(gp)))

    def has_subiterated(self, gp):
        return bool(self.lib.ref_has_subiterated(self.h, int(gp)))

    def write_restart(self, rid):
        self.lib.ref_write_restart(self.h, int(rid))

    def read_restart(self, rid):
        self.lib.ref_read_restart(self.h, int(rid))

    def output(self, gp, filename):
        self.lib.ref_output(self.h, int(gp), str(filename).encode())

    # inspection
    def elem_type(self):
        out = np.zeros(self.nelem, dtype=np.int32)
        self.lib.ref_get_elem_type(self.h, out.ctypes.data_as(_ip))
        return out

    def bmat(self):
        out = np.zeros((8, 6, 24))
        self.lib.ref_get_bmat(self.h, _d(out))
        return out

    def ctan_lin(self):
        out = np.zeros(36)
        self.lib.ref_get_ctan_lin(self.h, _d(out))
        return out

    def get_u(self, gp, which=1):
        out = np.zeros(self.nndim)
        ok = self.lib.ref_get_u(self.h, int(gp), int(which), _d(out))
        return out if ok else None

    def get_vars(self, gp, which=0):
        out = np.zeros(self.nvars)
        ok = self.lib.ref_get_vars(self.h, int(gp), int(which), _d(out))
        return out if ok else None

    # FE stages
    def set_displ_bc(self, eps, u=None):
        u = np.zeros(self.nndim) if u is None else np.ascontiguousarray(u, dtype=np.float64).copy()
        e = np.ascontiguousarray(eps, dtype=np.float64)
        self.lib.ref_set_displ_bc(self.h, _d(e), _d(u))
        return u

    def assembly_rhs(self, u, vars_old=None):
        b = np.zeros(self.nndim)
        nrm = self.lib.ref_assembly_rhs(self.h, _d(u), _vars_ptr(vars_old), _d(b))
        return b, float(nrm)

    def assembly_mat(self, u, vars_old=None):
        vals = np.zeros((self.nndim, 81))
        self.lib.ref_assembly_mat(self.h, _d(u), _vars_ptr(vars_old), _d(vals))
        return vals

    def ave_stress(self, u, vars_old=None):
        s = np.zeros(6)
        self.lib.ref_ave_stress(self.h, _d(u), _vars_ptr(vars_old), _d(s))
        return s

    def ave_strain(self, u):
        s = np.zeros(6)
        self.lib.ref_ave_strain(self.h, _d(u), _d(s))
        return s

    def vars_new(self, u, vars_old=None):
        vn = np.zeros(self.nvars)
        nl = self.lib.ref_vars_new(self.h, _d(u), _vars_ptr(vars_old), _d(vn))
        return vn, bool(nl)

    def newton(self, eps, u, vars_old=None):
        u = np.ascontiguousarray(u, dtype=np.float64).copy()
        e = np.ascontiguousarray(eps, dtype=np.float64)
        out = np.zeros(3, dtype=np.int32)
        self.lib.ref_newton(self.h, _d(e), _vars_ptr(vars_old), _d(u), out.ctypes.data_as(_ip))
        return u, dict(its=int(out[0]), solver_its=int(out[1]), converged=bool(out[2]))


# free functions -------------------------------------------------------------------------

def ell_cols(nx, ny, nz):
    lib = load()
    cols = np.zeros((3 * nx * ny * nz, 81), dtype=np.int32)
    lib.ref_ell_cols(nx, ny, nz, cols.ctypes.data_as(_ip))
    return cols


def elem_nodes(nx, ny, ex, ey, ez):
    lib = load()
    n = np.zeros(8, dtype=np.int32)
    lib.ref_elem_nodes(nx, ny, ex, ey, ez, n.ctypes.data_as(_ip))
    return n


def ell_add_one(nx, ny, nz, ex, ey, ez, Ae):
    lib = load()
    vals = np.zeros((3 * nx * ny * nz, 81))
    Ae = np.ascontiguousarray(Ae, dtype=np.float64)
    lib.ref_ell_add_one(nx, ny, nz, ex, ey, ez, _d(Ae), _d(vals))
    return vals


def ell_mvp(nx, ny, nz, vals, x):
    lib = load()
    y = np.zeros(3 * nx * ny * nz)
    lib.ref_ell_mvp(nx, ny, nz, _d(np.ascontiguousarray(vals)), _d(np.ascontiguousarray(x)), _d(y))
    return y


def ell_solve_cgpd(nx, ny, nz, vals, b):
    lib = load()
    x = np.zeros(3 * nx * ny * nz)
    err = C.c_double(0.0)
    its = lib.ref_ell_solve_cgpd(nx, ny, nz, _d(np.ascontiguousarray(vals)), _d(np.ascontiguousarray(b)), _d(x),
                                 C.byref(err))
    return x, int(its), float(err.value)


def _mat_args(m):
    return [C.c_int(int(m[0]))] + [C.c_double(float(v)) for v in m[1:6]]


def mat_stress(m, eps, vars_old=None):
    lib = load()
    s = np.zeros(6)
    lib.ref_mat_stress(*_mat_args(m), _d(np.ascontiguousarray(eps, dtype=np.float64)), _vars_ptr(vars_old), _d(s))
    return s


def mat_ctan(m, eps, vars_old=None):
    lib = load()
    c = np.zeros(36)
    lib.ref_mat_ctan(*_mat_args(m), _d(np.ascontiguousarray(eps, dtype=np.float64)), _vars_ptr(vars_old), _d(c))
    return c


def mat_evolute(m, eps, vars_old=None):
    lib = load()
    vn = np.zeros(7)
    nl = lib.ref_mat_evolute(*_mat_args(m), _d(np.ascontiguousarray(eps, dtype=np.float64)), _vars_ptr(vars_old),
                             _d(vn))
    return vn, bool(nl)


def mvp3(m, x):
    lib = load()
    y = np.zeros(3)
    lib.ref_mvp3(_d(np.ascontiguousarray(m, dtype=np.float64)), _d(np.ascontiguousarray(x, dtype=np.float64)), _d(y))
    return y


def omp_max_threads():
    return int(load(True).ref_omp_max_threads())
