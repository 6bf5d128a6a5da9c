"""One large RVE split into z-slabs over several GPUs (BASELINE.json configs[4]; SURVEY.md 8e "single oversized RVE").

The reference has no such mode -- its only parallelism is over independent Gauss points.  Here the node planes of one
nx x ny x nz RVE are divided into contiguous z-ranges, one per rank (one process per GPU).  Every rank runs the SAME
sm_100a kernels as the batched path on its local planes plus one halo plane towards each neighbour
(`micropp3x_slab_create`, include/micropp_b200_ext.h); the only differences are

  * before every SpMV the halo planes of the search direction p are fetched from the two neighbours
    (3 * nx * ny doubles each way): by default PULLED straight from the neighbour's vector over NVLink peer memory
    (CUDA IPC mapping, device-side epoch flags; `exchange="peer"`), or by NCCL send/recv (`exchange="nccl"`),
  * the three dot products of a DPCG iteration (p.Ap, then z.z and r.z), the residual norm of a Newton step and
    the six stress sums are slab-local sums that are summed over the ranks -- through peer-mapped mailboxes, in rank
    order, inside the tail kernel itself (peer mode), or by an NCCL all-reduce -- before their scalar "tail"
    (alpha, beta, convergence tests: the reference's logic of src/ell.cpp:93-119 and src/solve.cpp:43-47) runs on
    every rank -- so all ranks take identical decisions.

No other vector needs an exchange: x += alpha p keeps the halo entries of du (and hence u) consistent because the
halo entries of p are the neighbour's values and alpha is global.  Element layers that touch a cut are assembled by
both sides; each layer is averaged by exactly one rank.

`SlabRVE` can also hold several slabs in ONE process on one GPU (`world=None, nslabs=R`): the same kernels with plain
device pointers -- used to test the decomposition on a single GPU.

Host logic: the product path (`exchange="peer"`) is C++ -- micropp_b200/csrc/slab_host.cpp behind the C ABI
`micropp3x_slab_*` of include/micropp_b200_ext.h; this module only exchanges the CUDA IPC handles through the process
group (a C++ macro code would use MPI_Allgather) and forwards the calls.  `exchange="nccl"` keeps a Python-driven loop
over the same kernels with torch.distributed send/recv + all-reduce as the BASELINE transport it is measured against.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import Micropp3Params, default_params, load

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class SlotState(C.Structure):
    """mgpu_slot_state (include/mgpu.h)."""
    _fields_ = [("norm0", C.c_double), ("norm", C.c_double), ("rz", C.c_double), ("pAp", C.c_double),
                ("alpha", C.c_double), ("beta", C.c_double), ("pnorm0", C.c_double), ("pnorm", C.c_double),
                ("nr_its", C.c_int), ("solver_its", C.c_int), ("nr_active", C.c_int), ("converged", C.c_int),
                ("cg_its", C.c_int), ("cg_active", C.c_int), ("nl_flag", C.c_int), ("ticket", C.c_uint),
                ("cg_hist", C.c_void_p), ("cg_hist_k", C.c_int), ("pad_", C.c_int)]


def plane_range(nz: int, nslabs: int, s: int) -> tuple[int, int]:
    """Node planes [z0, z1) owned by slab s: contiguous, remainder to the low slabs (the rule the reference's MPI
    drivers use for Gauss points, test/multi-gpu-mpi.cpp:60)."""
    cnt = [nz // nslabs + (1 if nz % nslabs > r else 0) for r in range(nslabs)]
    z0 = sum(cnt[:s])
    return z0, z0 + cnt[s]


def exchange_planes(dist, rank: int, size: int, send_lo, send_hi, recv_lo, recv_hi):
    """One halo exchange with the z-neighbours: rank-1 gets `send_lo` and fills `recv_lo`, rank+1 gets `send_hi` and
    fills `recv_hi` (None where there is no neighbour).  All transfers go out as ONE batched send/recv group."""
    ops = []
    if rank + 1 < size:
        ops += [dist.P2POp(dist.isend, send_hi, rank + 1), dist.P2POp(dist.irecv, recv_hi, rank + 1)]
    if rank > 0:
        ops += [dist.P2POp(dist.isend, send_lo, rank - 1), dist.P2POp(dist.irecv, recv_lo, rank - 1)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class _DevArray:
    """Zero-copy view of library-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _bind(lib):
    if getattr(lib, "_slab_bound", False):
        return
    V = C.c_void_p
    sig = {
        "micropp3x_slab_create": (V, [C.POINTER(Micropp3Params), C.c_int, C.c_int, C.c_int]),
        "mgpu_destroy": (None, [V]), "mgpu_nn_pad": (C.c_int, [V]), "mgpu_sync": (None, [V]),
        "mgpu_launch_count": (C.c_ulonglong, [V]),
        "mgpu_bind_slots": (None, [V, C.c_int, _ip, _ip, _ip]), "mgpu_set_list": (None, [V, C.c_int, C.c_int, _ip]),
        "mgpu_set_slot_strain": (None, [V, C.c_int, _ip, _dp]),
        "mgpu_zero_u": (None, [V, C.c_int, C.c_int]), "mgpu_set_bc": (None, [V, C.c_int, C.c_int]),
        "mgpu_asm_rhs": (None, [V, C.c_int, C.c_int, C.c_int]), "mgpu_asm_mat": (None, [V, C.c_int, C.c_int, C.c_int]),
        "mgpu_cg_init": (None, [V, C.c_int, C.c_int, C.c_int]),
        "mgpu_cg_spmv_dot": (None, [V, C.c_int, C.c_int, C.c_int]),
        "mgpu_cg_update": (None, [V, C.c_int, C.c_int]), "mgpu_cg_pupdate": (None, [V, C.c_int, C.c_int]),
        "mgpu_cg_finish": (None, [V, C.c_int, C.c_int]),
        "mgpu_axpy_u": (None, [V, C.c_int, C.c_int]), "mgpu_ave_stress": (None, [V, C.c_int, C.c_int]),
        "mgpu_tail": (None, [V, C.c_int, C.c_int, C.c_int, C.c_int]),
        "mgpu_fetch_state": (None, [V, C.c_int, _ip, C.POINTER(SlotState)]),
        "mgpu_fetch_stress": (None, [V, C.c_int, _ip, _dp]),
        "mgpu_dev_ptr": (V, [V, C.c_int]), "mgpu_stream": (V, [V]), "mgpu_implicit": (C.c_int, [V]),
        "mgpu_stage_get_u": (None, [V, C.c_int, _dp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    lib.mgpu_slot_state_size.restype = C.c_int
    if lib.mgpu_slot_state_size() != C.sizeof(SlotState):   # mgpu_fetch_state writes whole structs into our mirror
        raise RuntimeError(f"slab.py: SlotState mirrors {C.sizeof(SlotState)} bytes, mgpu_slot_state has "
                           f"{lib.mgpu_slot_state_size()} (include/mgpu.h changed)")
    lib._slab_bound = True


class _Slab:
    """Device context of one z-slab (planes [z0, z1) + halos) and torch views of its exchange buffers."""

    def __init__(self, lib, cparams, dims, z0, z1, device):
        import torch
        self.lib, self.z0, self.z1 = lib, z0, z1
        nx, ny, nz = dims
        self.nxny = nx * ny
        self.halo_lo, self.halo_hi = int(z0 > 0), int(z1 < nz)
        self.koff = z0 - self.halo_lo
        self.nzl = z1 + self.halo_hi - self.koff
        self.ctx = C.c_void_p(lib.micropp3x_slab_create(C.byref(cparams), z0, z1, device))
        self.nn_pad = lib.mgpu_nn_pad(self.ctx)
        dev = torch.device("cuda", device)
        self.p = torch.as_tensor(_DevArray(lib.mgpu_dev_ptr(self.ctx, 3), 3 * self.nn_pad), device=dev).view(3, -1)
        self.red = torch.as_tensor(_DevArray(lib.mgpu_dev_ptr(self.ctx, 10), 8), device=dev)
        self.stream = torch.cuda.ExternalStream(int(lib.mgpu_stream(self.ctx)), device=dev)
        self.zero = (C.c_int * 1)(0)
        minus = (C.c_int * 1)(-1)
        lib.mgpu_bind_slots(self.ctx, 1, self.zero, minus, None)
        lib.mgpu_set_list(self.ctx, 0, 1, self.zero)

    def plane(self, k_local):
        """p on local plane k: a [3, nx*ny] view."""
        return self.p[:, k_local * self.nxny:(k_local + 1) * self.nxny]

    def state(self) -> SlotState:
        st = SlotState()
        self.lib.mgpu_fetch_state(self.ctx, 1, self.zero, C.byref(st))
        return st

    def close(self):
        if self.ctx:
            self.lib.mgpu_destroy(self.ctx)
            self.ctx = None


class SlabRVEPy:
    """PYTHON-DRIVEN variant (the NCCL baseline transport and the unfused peer kernels; see SlabRVE for the product path).

    One RVE, FE_ONE_WAY without history (elastic, or the first load step of any law), solved over z-slabs.

    world=(dist, rank, size): one slab per process, NCCL collectives.  world=None: `nslabs` slabs in this process on
    one GPU (test mode)."""

    L0 = 0  # device slot list used for every launch (one slot per slab)

    def __init__(self, params: dict, *, world=None, nslabs: int = 1, device: int = 0, cg_chunk: int = 8,
                 exchange: str = "peer"):
        import torch
        self.torch = torch
        self.lib = load()
        _bind(self.lib)
        p = default_params(**params)
        self.p = p
        self.dims = tuple(int(v) for v in p["size"])
        nz = self.dims[2]
        cp = Micropp3Params()
        cp.ngp = 1
        cp.size[:] = self.dims
        cp.type = int(p["type"])
        cp.geo_params[:] = [float(v) for v in p["geo_params"]]
        for i, m in enumerate(p["materials"][:3]):
            cp.mat_type[i] = int(m[0])
            cp.mat_E[i], cp.mat_nu[i], cp.mat_Ka[i], cp.mat_Sy[i], cp.mat_Xt[i] = [float(v) for v in m[1:6]]
        cp.nr_max_its, cp.nr_max_tol, cp.nr_rel_tol = int(p["nr_max_its"]), float(p["nr_max_tol"]), float(p["nr_rel_tol"])
        self.world = world
        self.cg_chunk = cg_chunk
        if world is not None:
            self.dist, self.rank, self.size = world
            ranges = [plane_range(nz, self.size, self.rank)]
        else:
            self.dist, self.rank, self.size = None, 0, nslabs
            ranges = [plane_range(nz, nslabs, s) for s in range(nslabs)]
        if any(z1 - z0 < 1 for z0, z1 in ranges):
            raise ValueError("more slabs than node planes")
        self.slabs = [_Slab(self.lib, cp, self.dims, z0, z1, device) for z0, z1 in ranges]
        self.device = torch.device("cuda", device)
        n3 = 3 * self.dims[0] * self.dims[1]
        if world is not None:  # contiguous send/recv staging: 3 components of one plane per message
            self.sbuf = [torch.empty(n3, dtype=torch.float64, device=self.device) for _ in range(2)]
            self.rbuf = [torch.empty(n3, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.exchanges = 0
        self.allreduces = 0
        # DPCG operator: 3 = implicit operator of an all-elastic RVE (no assembled matrix), 0 = the slab's own ELL matrix;
        # ONE operator for the whole RVE: every rank must take the same decision (all-reduce of the local flag)
        imp = int(all(self.lib.mgpu_implicit(s.ctx) for s in self.slabs))
        if world is not None:
            t = torch.tensor([imp], dtype=torch.int32, device=self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            imp = int(t.item())
        self.op = 3 if imp else 0
        self.elastic = all(int(m[0]) == 0 for m in p["materials"][:3])
        self.solves = 0
        if exchange != "nccl":
            raise ValueError("SlabRVEPy is the NCCL baseline transport; the peer-memory path is SlabRVE (C++ host)")
        self.exchange = exchange
        self._mapped = []

    # ------------------------------------------------------------------ communication
    def _allreduce(self, k: int):
        """Sum the first k slab-local sums over all slabs, result to every slab."""
        torch = self.torch
        self.allreduces += 1
        if self.world is not None:
            s = self.slabs[0]
            with torch.cuda.stream(s.stream):
                self.dist.all_reduce(s.red[:k])
            return
        torch.cuda.synchronize()
        tot = self.slabs[0].red[:k].clone()
        for s in self.slabs[1:]:  # fixed slab order => deterministic
            tot += s.red[:k]
        for s in self.slabs:
            s.red[:k].copy_(tot)
        torch.cuda.synchronize()

    def _exchange_p(self):
        """Halo planes of p <- the neighbour's adjacent owned plane."""
        torch = self.torch
        self.exchanges += 1
        if self.world is None:
            torch.cuda.synchronize()
            for a, b in zip(self.slabs[:-1], self.slabs[1:]):  # a below b
                b.plane(0).copy_(a.plane(a.nzl - 2))     # b's low halo  <- a's top owned plane
                a.plane(a.nzl - 1).copy_(b.plane(1))     # a's high halo <- b's bottom owned plane
            torch.cuda.synchronize()
            return
        s = self.slabs[0]
        with torch.cuda.stream(s.stream):
            if s.halo_hi:
                self.sbuf[1].view(3, -1).copy_(s.plane(s.nzl - 2))
            if s.halo_lo:
                self.sbuf[0].view(3, -1).copy_(s.plane(1))
            exchange_planes(self.dist, self.rank, self.size, self.sbuf[0], self.sbuf[1], self.rbuf[0], self.rbuf[1])
            if s.halo_hi:
                s.plane(s.nzl - 1).copy_(self.rbuf[1].view(3, -1))
            if s.halo_lo:
                s.plane(0).copy_(self.rbuf[0].view(3, -1))

    def _each(self, fn, *args):
        for s in self.slabs:
            fn(s.ctx, *args)

    def _reduced(self, kernel, kargs, k, tail_kind, tail_mode=0):
        """reducing kernel -> all-reduce of its k slab-local sums -> scalar tail on every slab"""
        self._each(kernel, *kargs)
        self._allreduce(k)
        self._each(self.lib.mgpu_tail, self.L0, 1, tail_kind, tail_mode)

    # ------------------------------------------------------------------ solver
    def cg_solve(self):
        lib, L0 = self.lib, self.L0
        self._reduced(lib.mgpu_cg_init, (L0, 1, self.op), 2, 1)
        while self.slabs[0].state().cg_active:
            for _ in range(self.cg_chunk):
                self._exchange_p()
                self._reduced(lib.mgpu_cg_spmv_dot, (L0, 1, self.op), 1, 2)
                self._reduced(lib.mgpu_cg_update, (L0, 1), 2, 3)
                self._each(lib.mgpu_cg_pupdate, L0, 1)
        self._each(lib.mgpu_cg_finish, L0, 1)   # the deferred x += alpha p of the last iteration

    def homogenize(self, eps) -> dict:
        """set_displ_bc -> Newton-Raphson (src/solve.cpp:29-82) -> averaged stress (src/average.cpp:58-82)."""
        lib, L0 = self.lib, self.L0
        if not self.elastic and self.solves > 0:
            raise RuntimeError("the z-slab mode keeps no history: a non-elastic RVE can be solved once (virgin state)")
        self.solves += 1
        e = np.ascontiguousarray(eps, dtype=np.float64)
        for s in self.slabs:
            lib.mgpu_set_slot_strain(s.ctx, 1, s.zero, e.ctypes.data_as(_dp))
        self._each(lib.mgpu_zero_u, L0, 1)
        self._each(lib.mgpu_set_bc, L0, 1)
        self._reduced(lib.mgpu_asm_rhs, (L0, 1, 0), 1, 0, 0)
        while self.slabs[0].state().nr_active:
            if self.op == 0:
                self._each(lib.mgpu_asm_mat, L0, 1, 0)
            self.cg_solve()
            self._each(lib.mgpu_axpy_u, L0, 1)
            self._reduced(lib.mgpu_asm_rhs, (L0, 1, 1), 1, 0, 1)
        self._reduced(lib.mgpu_ave_stress, (L0, 1), 6, 4)
        sig = np.zeros(6)
        s0 = self.slabs[0]
        lib.mgpu_fetch_stress(s0.ctx, 1, s0.zero, sig.ctypes.data_as(_dp))
        st = s0.state()
        return dict(stress=sig, newton_its=st.nr_its, cg_its=st.solver_its, converged=bool(st.converged),
                    norm=st.norm, norm0=st.norm0)

    def get_u(self) -> np.ndarray:
        """Displacements of the owned planes of this process's slabs, reference layout [node][3], global z order."""
        nx, ny, _ = self.dims
        out = []
        for s in self.slabs:
            buf = np.zeros(3 * nx * ny * s.nzl)
            self.lib.mgpu_stage_get_u(s.ctx, 0, buf.ctypes.data_as(_dp))
            u = buf.reshape(s.nzl, ny * nx, 3)
            out.append(u[s.z0 - s.koff:s.z1 - s.koff])
        return np.concatenate(out, axis=0).reshape(-1, 3)

    def launch_count(self) -> int:
        return sum(int(self.lib.mgpu_launch_count(s.ctx)) for s in self.slabs)

    def sync(self):
        self._each(self.lib.mgpu_sync)

    def peer_error(self) -> int:
        return 0

    def close(self):
        for s in self.slabs:
            s.close()
        self.slabs = []


# ------------------------------------------------------------------------------------------------ product path (C++ host)
class SlabHandle(C.Structure):
    """micropp3x_slab_handle (include/micropp_b200_ext.h)."""
    _fields_ = [("mail", C.c_char * 64), ("op", C.c_int), ("pad", C.c_int)]


def _bind_cxx(lib):
    if getattr(lib, "_slab_cxx_bound", False):
        return
    V = C.c_void_p
    PP = C.POINTER(Micropp3Params)
    sig = {
        "micropp3x_slab_new": (V, [PP, C.c_int, C.c_int, C.c_int]), "micropp3x_slab_free": (None, [V]),
        "micropp3x_slab_export": (None, [V, C.POINTER(SlabHandle)]),
        "micropp3x_slab_connect": (None, [V, C.POINTER(SlabHandle)]),
        "micropp3x_slab_connect_local": (None, [C.POINTER(V), C.c_int]),
        "micropp3x_slab_homogenize": (C.c_int, [V, _dp, _dp, _ip]),
        "micropp3x_slab_homogenize_local": (C.c_int, [C.POINTER(V), C.c_int, _dp, _dp, _ip]),
        "micropp3x_slab_planes": (None, [V, _ip, _ip]), "micropp3x_slab_get_u": (None, [V, _dp]),
        "micropp3x_slab_launch_count": (C.c_ulonglong, [V]), "micropp3x_slab_operator": (C.c_int, [V]),
        "micropp3x_slab_cg_history": (None, [V, C.c_int]), "micropp3x_slab_cg_history_read": (C.c_int, [V, _dp, C.c_int]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    lib._slab_cxx_bound = True


def _cparams(p):
    cp = Micropp3Params()
    cp.ngp = 1
    cp.size[:] = [int(v) for v in p["size"]]
    cp.type = int(p["type"])
    cp.geo_params[:] = [float(v) for v in p["geo_params"]]
    for i, m in enumerate(p["materials"][:3]):
        cp.mat_type[i] = int(m[0])
        cp.mat_E[i], cp.mat_nu[i], cp.mat_Ka[i], cp.mat_Sy[i], cp.mat_Xt[i] = [float(v) for v in m[1:6]]
    cp.nr_max_its, cp.nr_max_tol, cp.nr_rel_tol = int(p["nr_max_its"]), float(p["nr_max_tol"]), float(p["nr_rel_tol"])
    return cp


class SlabRVECxx:
    """One RVE over z-slabs, driven by the C++ host (micropp3x_slab_*): peer-memory exchange, fused reductions, CUDA-graph
    chunks of DPCG iterations.  world=(dist, rank, size): one slab per process; world=None: `nslabs` slabs in-process."""

    exchange = "peer"

    def __init__(self, params: dict, *, world=None, nslabs: int = 1, device: int = 0):
        self.lib = load()
        _bind_cxx(self.lib)
        p = default_params(**params)
        self.dims = tuple(int(v) for v in p["size"])
        cp = _cparams(p)
        self.world = world
        V = C.c_void_p
        if world is not None:
            self.dist, self.rank, self.size = world
            h = self.lib.micropp3x_slab_new(C.byref(cp), self.rank, self.size, device)
            if not h:
                raise ValueError("more slabs than node planes")
            self.slabs = [V(h)]
            mine = SlabHandle()
            self.lib.micropp3x_slab_export(self.slabs[0], C.byref(mine))
            allraw = [None] * self.size
            self.dist.all_gather_object(allraw, bytes(mine))          # the only collective: set-up
            arr = (SlabHandle * self.size)(*[SlabHandle.from_buffer_copy(b) for b in allraw])
            self.lib.micropp3x_slab_connect(self.slabs[0], arr)
            self.dist.barrier()
        else:
            self.dist, self.rank, self.size = None, 0, nslabs
            hs = [self.lib.micropp3x_slab_new(C.byref(cp), r, nslabs, device) for r in range(nslabs)]
            if not all(hs):
                raise ValueError("more slabs than node planes")
            self.slabs = [V(h) for h in hs]
            self._group = (V * nslabs)(*self.slabs)
            self.lib.micropp3x_slab_connect_local(self._group, nslabs)
        self.op = int(self.lib.micropp3x_slab_operator(self.slabs[0]))
        self.exchanges = 0      # halo pulls / cross-rank sums (derived from the iteration counts: the loop is device-side)
        self.allreduces = 0
        self._err = 0

    def homogenize(self, eps) -> dict:
        e = np.ascontiguousarray(eps, dtype=np.float64)
        sig = np.zeros(6)
        out = (C.c_int * 3)()
        if self.world is not None:
            rc = self.lib.micropp3x_slab_homogenize(self.slabs[0], e.ctypes.data_as(_dp), sig.ctypes.data_as(_dp), out)
        else:
            rc = self.lib.micropp3x_slab_homogenize_local(self._group, len(self.slabs), e.ctypes.data_as(_dp),
                                                          sig.ctypes.data_as(_dp), out)
        if rc in (-1, -2):
            raise RuntimeError("micropp3x_slab_homogenize refused the call (see stderr): rc = %d" % rc)
        self._err = 1 if rc == -3 else 0
        self.exchanges += int(out[1])
        self.allreduces += 2 * int(out[1]) + int(out[0]) + 2
        return dict(stress=sig, newton_its=int(out[0]), cg_its=int(out[1]), converged=bool(out[2]))

    def get_u(self) -> np.ndarray:
        """Displacements of the owned planes of this process's slabs, reference layout [node][3], global z order."""
        nx, ny, nz = self.dims
        parts = []
        for s in self.slabs:
            z0, z1 = C.c_int(), C.c_int()
            self.lib.micropp3x_slab_planes(s, C.byref(z0), C.byref(z1))
            lo, hi = int(z0.value > 0), int(z1.value < nz)
            nzl = z1.value + hi - (z0.value - lo)
            buf = np.zeros(3 * nx * ny * nzl)
            self.lib.micropp3x_slab_get_u(s, buf.ctypes.data_as(_dp))
            parts.append(buf.reshape(nzl, ny * nx, 3)[lo:lo + z1.value - z0.value])
        return np.concatenate(parts, axis=0).reshape(-1, 3)

    def launch_count(self) -> int:
        return sum(int(self.lib.micropp3x_slab_launch_count(s)) for s in self.slabs)

    def cg_history(self, k):
        """Test instrument: record |z| at the head of the first k DPCG iterations of the latest solve."""
        for s in self.slabs:
            self.lib.micropp3x_slab_cg_history(s, int(k))

    def cg_history_read(self, k, slab=0):
        out = np.zeros(k)
        n = self.lib.micropp3x_slab_cg_history_read(self.slabs[slab], out.ctypes.data_as(_dp), int(k))
        return out[:n]

    def peer_error(self) -> int:
        return self._err

    def close(self):
        if self.world is not None and self.slabs:
            self.dist.barrier()          # nobody unmaps memory a neighbour may still read
        for s in self.slabs:
            self.lib.micropp3x_slab_free(s)
        self.slabs = []


def SlabRVE(params: dict, *, world=None, nslabs: int = 1, device: int = 0, cg_chunk: int = 8, exchange: str = "peer"):
    """One RVE solved over z-slabs.  exchange="peer" (default): the product path, host logic in C++ (SlabRVECxx);
    exchange="nccl": the Python-driven baseline transport (SlabRVEPy)."""
    if exchange == "peer":
        return SlabRVECxx(params, world=world, nslabs=nslabs, device=device)
    return SlabRVEPy(params, world=world, nslabs=nslabs, device=device, cg_chunk=cg_chunk, exchange=exchange)
