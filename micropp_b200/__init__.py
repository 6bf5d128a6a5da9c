"""micropp_b200 -- B200-native RVE homogenization (drop-in for gagiuntoli/Micropp's hot path).

The product is ``libmicropp_b200.so``: hand-written sm_100a CUDA kernels behind the reference's own
C ABI (``include/micropp_c.h``) and C++ class (``include/micropp.hpp``).  This Python module is only a
ctypes binding of that C ABI, used by the tests and ``bench.py``; it mirrors the method names of the
reference class.  There is NO CPU fallback: loading fails loudly if the library is missing, and
constructing a solver aborts if no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["MICROPP_B200_LIB"]) if os.environ.get("MICROPP_B200_LIB") else PKG / "libmicropp_b200.so"

# enums of include/types.hpp and include/material_base.h
MIC = dict(HOMOGENEOUS=0, SPHERE=1, LAYER_Y=2, CILI_FIB_X=3, CILI_FIB_Z=4, CILI_FIB_XZ=5, QUAD_FIB_XYZ=6,
           QUAD_FIB_XZ=7, QUAD_FIB_XZ_BROKEN_X=8, SPHERES=9, MIC3D_8=10, FIBS_20_ORDER=11, FIBS_20_DISORDER=12)
FE_LINEAR, FE_ONE_WAY, FE_FULL, MIX_RULE_CHAMIS = 0, 1, 2, 3
ELASTIC, PLASTIC, DAMAGE = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class MaterialBase(C.Structure):
    """struct material_base (include/material_base.h)."""
    _fields_ = [("E", C.c_double), ("nu", C.c_double), ("Ka", C.c_double), ("Sy", C.c_double),
                ("k", C.c_double), ("mu", C.c_double), ("lam", C.c_double), ("Xt", C.c_double), ("type", C.c_int)]


class Micropp3Handle(C.Structure):
    """struct micropp3 (include/micropp_c.h)."""
    _fields_ = [("ptr", C.c_void_p)]


class Micropp3Params(C.Structure):
    """struct micropp3_params (include/micropp_b200_ext.h)."""
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
    """Defaults of micropp_params_t (include/types.hpp)."""
    p = dict(ngp=1, size=(5, 5, 5), type=0, geo_params=(0.1, 0.1, 0.1, 0.1),
             materials=[(0, 1.0e7, 0.3, 0.0, 0.0, 0.0)] * 3,  # (type, E, nu, Ka, Sy, Xt)
             coupling=None, subiterations=False, nsubiterations=10, mpi_rank=0, nr_max_its=4, nr_max_tol=1.0e-10,
             nr_rel_tol=1.0e-3, calc_ctan_lin=True, use_A0=False, its_with_A0=1, lin_stress=True, write_log=False)
    p.update(kw)
    return p


_lib = None


def load():
    """Load libmicropp_b200.so (build it with ``python -m micropp_b200.build``).  Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: build the CUDA extension first "
                           f"(`python -m micropp_b200.build`); micropp_b200 has no CPU fallback")
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL)  # keep micropp<3> symbols apart from any reference build in-process
    H = C.POINTER(Micropp3Handle)
    sig = {
        "micropp3_new": (None, [H, C.c_int, _ip, C.c_int, _dp, C.POINTER(MaterialBase), _ip, C.c_int, C.c_int]),
        "micropp3_new_ext": (None, [H, C.POINTER(Micropp3Params)]),
        "micropp3_free": (None, [H]),
        "micropp3_set_strain": (None, [H, C.c_int, _dp]),
        "micropp3_get_stress": (None, [H, C.c_int, _dp]),
        "micropp3_get_ctan": (None, [H, C.c_int, _dp]),
        "micropp3_homogenize": (None, [H]),
        "micropp3_homogenize_linear": (None, [H]),
        "micropp3_update_vars": (None, [H]),
        "micropp3_is_non_linear": (C.c_bool, [H, C.c_int]),
        "micropp3_get_cost": (C.c_int, [H, C.c_int]),
        "micropp3_has_converged": (C.c_bool, [H, C.c_int]),
        "micropp3_has_subiterated": (C.c_bool, [H, C.c_int]),
        "micropp3_get_non_linear_gps": (C.c_int, [H]),
        "micropp3_write_restart": (None, [H, C.c_int]),
        "micropp3_read_restart": (None, [H, C.c_int]),
        "micropp3_print_info": (None, [H]),
        "micropp3_output": (None, [H, C.c_int, C.c_char_p]),
        "micropp3_output2": (None, [H, C.c_int, C.c_int, C.c_int]),
        "micropp3_set_strains": (None, [H, _dp]),
        "micropp3_get_stresses": (None, [H, _dp]),
        "micropp3_get_ctans": (None, [H, _dp]),
        "micropp3x_nelem": (C.c_int, [H]),
        "micropp3x_nndim": (C.c_int, [H]),
        "micropp3x_wave_size": (C.c_int, [H]), "micropp3x_implicit_rows": (C.c_int, [H]),
        "micropp3x_implicit_kernel": (C.c_int, [H]),
        "micropp3x_resident_info": (C.c_int, [H, _ip]), "micropp3x_prof_resident_ms": (C.c_double, [H, C.c_int]),
        "micropp3x_apply_operator": (C.c_double, [H, _dp, _dp, C.c_int, C.c_int]),
        "micropp3x_get_elem_type": (None, [H, _ip]),
        "micropp3x_get_bmat": (None, [H, _dp]),
        "micropp3x_get_ctan_lin": (None, [H, _dp]),
        "micropp3x_get_u": (C.c_int, [H, C.c_int, C.c_int, _dp]),
        "micropp3x_get_vars": (C.c_int, [H, C.c_int, C.c_int, _dp]),
        "micropp3x_set_displ_bc": (None, [H, _dp, _dp]),
        "micropp3x_assembly_rhs": (C.c_double, [H, _dp, _dp, _dp]),
        "micropp3x_assembly_mat": (None, [H, _dp, _dp, _dp]),
        "micropp3x_newton": (None, [H, _dp, _dp, _dp, _ip]),
        "micropp3x_ave_stress": (None, [H, _dp, _dp, _dp]),
        "micropp3x_vars_new": (C.c_int, [H, _dp, _dp, _dp]),
        "micropp3x_ell_cols": (None, [C.c_int, C.c_int, C.c_int, _ip]),
        "micropp3x_ell_mvp": (None, [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]),
        "micropp3x_ell_solve_cgpd": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp]),
        "micropp3x_elem_nodes": (None, [C.c_int] * 5 + [_ip]),
        "micropp3x_elem_colour": (C.c_int, [C.c_int] * 3),
        "micropp3x_prof_enable": (None, [H, C.c_int]),
        "micropp3x_prof_read": (None, [H, _dp, C.c_int]), "micropp3x_hybrid_available": (C.c_int, [H]),
        "micropp3x_cg_history": (None, [H, C.c_int]), "micropp3x_cg_history_read": (C.c_int, [H, C.c_int, _dp, C.c_int]),
        "micropp3x_last_homogenize_ms": (C.c_double, [H]),
        "micropp3x_launch_count": (C.c_ulonglong, [H]),
        "micropp3x_bench_spmv": (C.c_double, [H, C.c_int, C.c_int]),
        "micropp3x_bench_imp_spmv": (C.c_double, [H, C.c_int, C.c_int, C.c_int]),
        "micropp3x_bench_resident": (C.c_double, [H, C.c_int, C.c_int, C.c_int]),
        "micropp3x_resident_timeline": (None, [H, C.c_int, C.POINTER(C.c_longlong)]),
        "material_set": (None, [C.POINTER(MaterialBase), C.c_int] + [C.c_double] * 5),
        "mgpu_device_count": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib


def device_count() -> int:
    return int(load().mgpu_device_count())


def _d(a):
    return a.ctypes.data_as(_dp)


def _opt(a):
    return None if a is None else _d(a)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Micropp3:
    """``micropp<3>`` through the C ABI.  Method names follow include/micropp.hpp of the reference."""

    def __init__(self, params: dict | None = None, *, c_api: bool = False, **kw):
        self.lib = load()
        p = default_params(**(params or {}))
        p.update(kw)
        self.p = p
        self.ngp = int(p["ngp"])
        self.nx, self.ny, self.nz = [int(v) for v in p["size"]]
        self.nn = self.nx * self.ny * self.nz
        self.nndim = 3 * self.nn
        self.nelem = (self.nx - 1) * (self.ny - 1) * (self.nz - 1)
        self.nvars = self.nelem * 56
        self.h = Micropp3Handle()
        self._keep = []
        if c_api:
            # the reference's own constructor signature (include/micropp_c.h): most parameters are fixed by it
            mats = (MaterialBase * 3)()
            for i, m in enumerate(p["materials"][:3]):
                self.lib.material_set(C.byref(mats[i]), int(m[0]), *[float(v) for v in m[1:6]])
            size = (C.c_int * 3)(self.nx, self.ny, self.nz)
            geo = (C.c_double * 4)(*[float(v) for v in p["geo_params"]])
            cpl = np.ascontiguousarray(p["coupling"] if p["coupling"] is not None else [FE_ONE_WAY] * self.ngp,
                                       dtype=np.int32)
            self.lib.micropp3_new(C.byref(self.h), self.ngp, size, int(p["type"]), geo, mats,
                                  cpl.ctypes.data_as(_ip), int(p["nsubiterations"]), int(p["mpi_rank"]))
        else:
            s = Micropp3Params()
            s.ngp = self.ngp
            s.size[:] = [self.nx, self.ny, self.nz]
            s.type = int(p["type"])
            s.geo_params[:] = [float(v) for v in p["geo_params"]]
            for i, m in enumerate(p["materials"][:3]):
                s.mat_type[i] = int(m[0])
                s.mat_E[i], s.mat_nu[i], s.mat_Ka[i], s.mat_Sy[i], s.mat_Xt[i] = [float(v) for v in m[1:6]]
            if p.get("coupling") is not None:
                arr = np.ascontiguousarray(p["coupling"], dtype=np.int32)
                self._keep.append(arr)
                s.coupling = arr.ctypes.data_as(_ip)
            for k in ("subiterations", "nsubiterations", "mpi_rank", "nr_max_its", "calc_ctan_lin", "use_A0",
                      "its_with_A0", "lin_stress", "write_log"):
                setattr(s, k, int(p[k]))
            s.nr_max_tol = float(p["nr_max_tol"])
            s.nr_rel_tol = float(p["nr_rel_tol"])
            self.lib.micropp3_new_ext(C.byref(self.h), C.byref(s))

    def close(self):
        if self.h.ptr:
            self.lib.micropp3_free(C.byref(self.h))
            self.h.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- macro-scale coupling -----------------------------------------------------------------
    def set_strain(self, gp, eps):
        self.lib.micropp3_set_strain(C.byref(self.h), int(gp), _d(_f64(eps)))

    def set_strains(self, eps):
        e = _f64(eps)
        assert e.size == 6 * self.ngp
        self.lib.micropp3_set_strains(C.byref(self.h), _d(e))

    def get_stress(self, gp):
        out = np.zeros(6)
        self.lib.micropp3_get_stress(C.byref(self.h), int(gp), _d(out))
        return out

    def get_stresses(self):
        out = np.zeros((self.ngp, 6))
        self.lib.micropp3_get_stresses(C.byref(self.h), _d(out))
        return out

    def get_ctan(self, gp):
        out = np.zeros(36)
        self.lib.micropp3_get_ctan(C.byref(self.h), int(gp), _d(out))
        return out

    def get_ctans(self):
        out = np.zeros((self.ngp, 36))
        self.lib.micropp3_get_ctans(C.byref(self.h), _d(out))
        return out

    def homogenize(self):
        self.lib.micropp3_homogenize(C.byref(self.h))

    def homogenize_linear(self):
        self.lib.micropp3_homogenize_linear(C.byref(self.h))

    def update_vars(self):
        self.lib.micropp3_update_vars(C.byref(self.h))

    def is_non_linear(self, gp):
        return int(self.lib.micropp3_is_non_linear(C.byref(self.h), int(gp)))

    def get_non_linear_gps(self):
        return int(self.lib.micropp3_get_non_linear_gps(C.byref(self.h)))

    def get_cost(self, gp):
        return int(self.lib.micropp3_get_cost(C.byref(self.h), int(gp)))

    def has_converged(self, gp):
        return bool(self.lib.micropp3_has_converged(C.byref(self.h), int(gp)))

    def has_subiterated(self, gp):
        return bool(self.lib.micropp3_has_subiterated(C.byref(self.h), int(gp)))

    def write_restart(self, rid):
        self.lib.micropp3_write_restart(C.byref(self.h), int(rid))

    def read_restart(self, rid):
        self.lib.micropp3_read_restart(C.byref(self.h), int(rid))

    def print_info(self):
        self.lib.micropp3_print_info(C.byref(self.h))

    def output(self, gp, filename):
        """Write <filename>.vtu for one Gauss point (src/output.cpp:30-41)."""
        self.lib.micropp3_output(C.byref(self.h), int(gp), str(filename).encode())

    def output2(self, gp, elem_global, time_step):
        """Write micropp-<elem_global>-<time_step>.vtu into the working directory (src/output.cpp:44-69)."""
        self.lib.micropp3_output2(C.byref(self.h), int(gp), int(elem_global), int(time_step))

    # ---- inspection -----------------------------------------------------------------------------
    def wave_size(self):
        return int(self.lib.micropp3x_wave_size(C.byref(self.h)))

    def apply_operator(self, p, op=0, kernel=-1):
        """(A p, p.Ap) with the Jacobian at u = 0 through DPCG operator `op` (0 assembled ELL, 3 implicit)."""
        p = _f64(p)
        out = np.zeros(self.nndim)
        pap = self.lib.micropp3x_apply_operator(C.byref(self.h), _d(p), _d(out), int(op), int(kernel))
        return out, float(pap)

    def implicit_kernel(self):
        """-1: assembled ELL matrices; else the implicit SpMV kernel: 0 simple, 1 tiled (cp.async), 2 tiled (TMA)."""
        return int(self.lib.micropp3x_implicit_kernel(C.byref(self.h)))

    def implicit_rows(self):
        return int(self.lib.micropp3x_implicit_rows(C.byref(self.h)))

    def resident_info(self):
        """Cluster-resident DPCG (whole solve in one launch, one thread-block cluster per RVE): None when the
        three-kernel loop runs, else dict(cs, py, pz, tn, threads, smem, fixcap, clusters)."""
        meta = np.zeros(8, dtype=np.int32)
        if int(self.lib.micropp3x_resident_info(C.byref(self.h), meta.ctypes.data_as(_ip))) == 0:
            return None
        return dict(zip(("cs", "py", "pz", "tn", "threads", "smem", "fixcap", "clusters"), (int(v) for v in meta)))

    def prof_resident_ms(self, reset=True):
        return float(self.lib.micropp3x_prof_resident_ms(C.byref(self.h), int(bool(reset))))

    def elem_type(self):
        out = np.zeros(max(self.nelem, 1), dtype=np.int32)
        self.lib.micropp3x_get_elem_type(C.byref(self.h), out.ctypes.data_as(_ip))
        return out[:self.nelem]

    def bmat(self):
        out = np.zeros((8, 6, 24))
        self.lib.micropp3x_get_bmat(C.byref(self.h), _d(out))
        return out

    def ctan_lin(self):
        out = np.zeros(36)
        self.lib.micropp3x_get_ctan_lin(C.byref(self.h), _d(out))
        return out

    def get_u(self, gp, which=1):
        out = np.zeros(self.nndim)
        ok = self.lib.micropp3x_get_u(C.byref(self.h), int(gp), int(which), _d(out))
        return out if ok else None

    def get_vars(self, gp, which=0):
        out = np.zeros(self.nvars)
        ok = self.lib.micropp3x_get_vars(C.byref(self.h), int(gp), int(which), _d(out))
        return out if ok else None

    # ---- FE stages (protected members of the reference class) ------------------------------------
    def set_displ_bc(self, eps, u=None):
        u = np.zeros(self.nndim) if u is None else _f64(u).copy()
        self.lib.micropp3x_set_displ_bc(C.byref(self.h), _d(_f64(eps)), _d(u))
        return u

    def assembly_rhs(self, u, vars_old=None):
        b = np.zeros(self.nndim)
        nrm = self.lib.micropp3x_assembly_rhs(C.byref(self.h), _d(_f64(u)), _opt(vars_old), _d(b))
        return b, float(nrm)

    def assembly_mat(self, u, vars_old=None):
        vals = np.zeros((self.nndim, 81))
        self.lib.micropp3x_assembly_mat(C.byref(self.h), _d(_f64(u)), _opt(vars_old), _d(vals))
        return vals

    def newton(self, eps, u, vars_old=None):
        u = _f64(u).copy()
        out = np.zeros(3, dtype=np.int32)
        self.lib.micropp3x_newton(C.byref(self.h), _d(_f64(eps)), _opt(vars_old), _d(u), out.ctypes.data_as(_ip))
        return u, dict(its=int(out[0]), solver_its=int(out[1]), converged=bool(out[2]))

    def ave_stress(self, u, vars_old=None):
        s = np.zeros(6)
        self.lib.micropp3x_ave_stress(C.byref(self.h), _d(_f64(u)), _opt(vars_old), _d(s))
        return s

    def vars_new(self, u, vars_old=None):
        vn = np.zeros(self.nvars)
        nl = self.lib.micropp3x_vars_new(C.byref(self.h), _d(_f64(u)), _opt(vars_old), _d(vn))
        return vn, bool(nl)

    # ---- measurement --------------------------------------------------------------------------------
    def prof_enable(self, on=True):
        self.lib.micropp3x_prof_enable(C.byref(self.h), int(bool(on)))

    def prof_read(self, reset=True):
        out = np.zeros(9)
        self.lib.micropp3x_prof_read(C.byref(self.h), _d(out), int(bool(reset)))
        return dict(spmv_ms=out[0], spmv_launches=int(out[1]), spmv_slot_apps=int(out[2]), asm_mat_ms=out[3],
                    asm_rhs_ms=out[4], cg_vec_ms=out[5], hybrid_spmv_ms=out[6], hybrid_slot_apps=int(out[7]),
                    hybrid_row_apps=int(out[8]))

    def cg_history(self, k):
        """Test instrument: record |z| at the head of the first k DPCG iterations of every slot's latest solve."""
        self.lib.micropp3x_cg_history(C.byref(self.h), int(k))

    def cg_history_read(self, slot, k):
        out = np.zeros(k)
        n = self.lib.micropp3x_cg_history_read(C.byref(self.h), int(slot), _d(out), int(k))
        return out[:n]

    def hybrid_available(self):
        """True when RVEs with a damage / plastic phase may take the hybrid operator (implicit elastic row blocks +
        explicit rows only where an element has left its linear regime)."""
        return bool(self.lib.micropp3x_hybrid_available(C.byref(self.h)))

    def last_homogenize_ms(self):
        return float(self.lib.micropp3x_last_homogenize_ms(C.byref(self.h)))

    def launch_count(self):
        return int(self.lib.micropp3x_launch_count(C.byref(self.h)))

    def bench_spmv(self, nslots, iters=20):
        return float(self.lib.micropp3x_bench_spmv(C.byref(self.h), int(nslots), int(iters)))

    def bench_resident(self, nslots, reps=3, dbg=4):
        """ms per isolated launch of the cluster-resident DPCG kernel over nslots RVEs (dbg bit 4: exactly 60 iterations)."""
        return float(self.lib.micropp3x_bench_resident(C.byref(self.h), int(nslots), int(reps), int(dbg)))

    def resident_timeline(self, slot=0):
        """[8 CTAs][16 warps][8 phases] cycle counters of the last bench_resident(..., dbg | 256) run."""
        out = np.zeros(1024, dtype=np.int64)
        self.lib.micropp3x_resident_timeline(C.byref(self.h), int(slot), out.ctypes.data_as(C.POINTER(C.c_longlong)))
        return out.reshape(8, 16, 8)

    def bench_imp_spmv(self, nslots, iters=20, kern=2):
        """ms per application of the implicit elastic operator on `nslots` RVEs (kern: 2 context default, 10+v TMA variant v)."""
        return float(self.lib.micropp3x_bench_imp_spmv(C.byref(self.h), int(nslots), int(iters), int(kern)))


# ---- free functions of the ELL API ----------------------------------------------------------------------

def ell_cols(nx, ny, nz):
    cols = np.zeros((3 * nx * ny * nz, 81), dtype=np.int32)
    load().micropp3x_ell_cols(nx, ny, nz, cols.ctypes.data_as(_ip))
    return cols


def elem_nodes(nx, ny, ex, ey, ez):
    n = np.zeros(8, dtype=np.int32)
    load().micropp3x_elem_nodes(nx, ny, ex, ey, ez, n.ctypes.data_as(_ip))
    return n


def elem_colour(ex, ey, ez):
    return int(load().micropp3x_elem_colour(ex, ey, ez))


def ell_mvp(nx, ny, nz, vals, x):
    y = np.zeros(3 * nx * ny * nz)
    load().micropp3x_ell_mvp(nx, ny, nz, _d(_f64(vals)), _d(_f64(x)), _d(y))
    return y


def ell_solve_cgpd(nx, ny, nz, vals, b):
    x = np.zeros(3 * nx * ny * nz)
    err = C.c_double(0.0)
    its = load().micropp3x_ell_solve_cgpd(nx, ny, nz, _d(_f64(vals)), _d(_f64(b)), _d(x),
                                          C.cast(C.byref(err), _dp))
    return x, int(its), float(err.value)
