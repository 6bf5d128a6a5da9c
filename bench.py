#!/usr/bin/env python
"""bench.py -- homogenized Gauss points / second of the RVE-homogenization hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload elastic30|damage50] [--ngp G]
    python bench.py --impl reference ...      # the reference's own CPU/OpenMP path on the host cores

One "step" = one `homogenize()` pass over the whole resident batch of RVEs (one per macro Gauss point)
through the reference-facing C ABI of libmicropp_b200.so.  Workloads (BASELINE.json `configs`):

  elastic30 (default; configs[1])  1024 GPs per GPU, 30^3-node RVE, sphere r=0.2, elastic E=1e7/1e8 nu=.3,
                                   random strains U[-1e-3,1e-3]^6 (seed 1234), lin_stress=false, FE_ONE_WAY.
  damage50  (configs[2], per-GPU   512 GPs per GPU, 50^3-node RVE, sphere r=0.2, damage matrix E=1e7 nu=.3 Xt=1e5 +
             shard of 4096/8)      elastic sphere E=3e7, nr_max_its=12, load path eps_11 = s_g*0.1*t, t=k*0.015,
                                   s_g~U[0.5,1.5] (seed 1234).  Steps 0..5 of the path are run (with update_vars) as
                                   untimed preparation; every warm-up / timed step is load step 6 from that state
                                   (update_vars is not called in between, so every step does identical work).

Scaling is weak: every rank owns its own GPs (independent RVEs -- no data-path collective).

One invocation also measures, after the headline workload and inside the same JSON line (`workloads`), bounded
shards of the other BASELINE configs on the same GPUs: `damage50` and `plastic40` (fewer Gauss points per GPU, stated
in each entry; same keys as the headline: value, e2e, roofline, cpu_baseline) and `slab200` (ONE 200^3 elastic RVE over
z-slabs on all N GPUs: ms per solve, DPCG iterations).  `--no-extra` skips them; `--workload X` makes X the headline.
`--full-path` times the whole load path (every load step with update_vars) instead of a repeated single step.

JSON keys beyond the base contract:
  value      GP/s with the device-event time of homogenize() (CUDA events on the library's stream, max over ranks)
  e2e        GP/s with host buffers through the C ABI: set_strains(host) + homogenize + get_stresses(host), wall clock
  roofline   the dominant kernel k_spmv_dot (DPCG SpMV fused with p.Ap): CUDA events around every launch of an
             instrumented repeat of the timed steps (per-kernel events need plain stream launches; the timed steps
             themselves run one CUDA graph per Newton step).  Algorithmic bytes per RVE application = 1992 B *
             interior nodes (243 FP64 values read once, p read + Ap written; boundary rows are identity rows and are
             neither stored nor read).  SURVEY's 664 B/row figure is reported as `achieved_664` for reference.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/libmicropp_ref_omp.so, g++ -O3 -fopenmp) on the box's host
             cores, same workload, bounded sample of GPs (rank 0, N=1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

EL = lambda E, nu=0.3: (0, E, nu, 0.0, 0.0, 0.0)
DM = lambda E, nu, Xt: (2, E, nu, 0.0, 0.0, Xt)
PL = lambda E, nu, Ka, Sy: (1, E, nu, Ka, Sy, 0.0)

WORKLOADS = {
    "elastic30": dict(n=30, ngp=1024, prep_steps=0,
                      params=dict(type=1, geo_params=(0.2, 0.0, 0.0, 0.0), materials=[EL(1e7), EL(1e8), EL(1e7)],
                                  lin_stress=False, calc_ctan_lin=False),
                      label="configs[1]: 1024 GPs/GPU, 30^3-node RVE, sphere r=0.2, elastic contrast 10, FE_ONE_WAY"),
    "damage50": dict(n=50, ngp=512, prep_steps=6,
                     params=dict(type=1, geo_params=(0.2, 0.0, 0.0, 0.0),
                                 materials=[DM(1e7, 0.3, 1e5), EL(3e7), EL(3e7)],
                                 lin_stress=False, calc_ctan_lin=False, nr_max_its=12),
                     label="configs[2] per-GPU shard: 512 GPs/GPU, 50^3-node RVE, sphere r=0.2, damage matrix + "
                           "elastic sphere, load step 6/10 from the converged state of step 5"),
    # configs[3].  SURVEY 8d takes the materials of test/test3d_4.cpp:65-67, whose plastic phase (E = Sy = 1e3) cannot
    # yield below a strain of order one; the bench uses the reference's golden plastic material instead
    # (test/benchmark-plastic.cpp:78: E=3e7 nu=.25 Ka=1e7 Sy=1e5) so that state variables do evolve.
    "plastic40": dict(n=40, ngp=256, prep_steps=6,
                      params=dict(type=2, geo_params=(0.5, 0.0, 0.0, 0.0),
                                  materials=[EL(3e7, 0.25), PL(3e7, 0.25, 1e7, 1e5), EL(3e7, 0.25)],
                                  lin_stress=False, calc_ctan_lin=False, nr_max_its=12),
                      label="configs[3] per-GPU shard: 256 GPs/GPU, 40^3-node RVE, layers in y (width 0.5), elastic + "
                            "J2-plastic layer with state-variable update, load step 6/10 from the converged state of "
                            "step 5"),
}


def strains_for(workload: str, ngp: int, rank: int, step: int) -> np.ndarray:
    """Synthetic macro strains of GP batch `rank` (seeded: identical for every implementation)."""
    rng = np.random.default_rng(1234 + 7919 * rank)
    if workload == "elastic30":
        return rng.uniform(-1e-3, 1e-3, (ngp, 6))
    s = rng.uniform(0.5, 1.5, ngp)
    e = np.zeros((ngp, 6))
    if workload == "plastic40":
        e[:, 2] = s * 0.1 * 0.015 * step   # test/test3d_4.cpp:54 loads direction 2
    else:
        e[:, 0] = s * 0.1 * 0.015 * step
    return e


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i] == "Active"})
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "samples": len(sm), "reasons": reasons}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            for k in ("hbm_gbs", "hbm_gbps", "hbm_GBs"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


FP64_PEAK_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12   # 148 SMs x 64 DFMA/clk x 2 flop at the 1965 MHz boost clock (nominal)


def spmv_imp_bytes_per_rve(n: int) -> float:
    """Implicit operator of an all-elastic RVE: no matrix stream at all.  Per node 3 p values are read and (interior
    nodes) 3 Ap values written; the table of distinct row blocks (a few hundred KB) stays in L1/shared memory."""
    return 24.0 * n ** 3 + 24.0 * (n - 2) ** 3


def spmv_bytes_per_rve(n: int) -> tuple[float, float]:
    """(algorithmic bytes the kernel must move, SURVEY's 664 B/row figure) for one SpMV of one n^3 RVE.

    The ELL storage keeps interior rows only (boundary rows are identity rows with p = 0): per interior node 243 FP64
    values are read once, 3 p values read and 3 Ap values written: 1992 B * (n-2)^3.  p re-reads by neighbours are
    assumed cached."""
    nn, nint = n ** 3, (n - 2) ** 3
    return 1992.0 * nint, 664.0 * 3 * nn


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(name, wl, sample_ngp: int, steps: int, warmup: int, restart_from=None):
    """The reference's own CPU path (oracle/_ref, OpenMP over GPs) on a bounded sample of the workload.

    restart_from=(directory, id): the internal state at the start of the timed load step is read from a restart file in
    the reference's own format (src/output.cpp:217-262) instead of being recomputed on the CPU -- the preparation steps
    of damage50 / plastic40 take minutes there."""
    from oracle import refpy
    if not refpy.available(omp=True):
        raise RuntimeError("oracle/_ref/libmicropp_ref_omp.so missing: run `make -C oracle ref` where "
                           "/root/reference exists (the file travels with the gpurun snapshot)")
    cores = host_cores()
    # all host cores, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1 for its workers)
    os.environ["OMP_NUM_THREADS"] = os.environ.get("MICROPP_REF_THREADS", str(cores))
    n = wl["n"]
    p = refpy.default_params(size=(n, n, n), ngp=sample_ngp, **wl["params"])
    cwd = os.getcwd()
    if restart_from is not None:
        os.chdir(restart_from[0])
    try:
        r = refpy.RefMicropp(p, omp=True)

        def step(k):
            e = strains_for(name, sample_ngp, 0, k)
            for g in range(sample_ngp):
                r.set_strain(g, e[g])
            r.homogenize()
            return np.array([r.get_stress(g) for g in range(sample_ngp)])

        if restart_from is not None:
            r.read_restart(restart_from[1])
        else:
            for k in range(wl["prep_steps"]):
                step(k)
                r.update_vars()
        kk = wl["prep_steps"]
        for _ in range(warmup):
            step(kk)
        t0 = time.perf_counter()
        for _ in range(steps):
            sig = step(kk)
        dt = time.perf_counter() - t0
        cost = [r.get_cost(g) for g in range(sample_ngp)]
        r.close()
    finally:
        os.chdir(cwd)
    threads = min(int(os.environ["OMP_NUM_THREADS"]), sample_ngp)
    return dict(value=sample_ngp * steps / dt, ms_per_step=dt / steps * 1e3, cores=threads, stress=sig,
                sample=f"{sample_ngp} of the workload's GPs per step, {steps} timed step(s) after {warmup} warm-up; "
                       f"mean CG its/GP {np.mean(cost):.1f}; OMP_NUM_THREADS={os.environ['OMP_NUM_THREADS']} "
                       f"({threads} busy: one RVE per thread)"
                       + ("; state at the start of the timed load step read from a restart file in the reference's "
                          "format (written by the product after the preparation steps)" if restart_from else ""))


# ------------------------------------------------------------------------------------------------ B200 arm
class Env:
    """process-group plumbing (torch.distributed over NCCL): barrier + reductions of the timings, nothing else"""

    def __init__(self, torch, dist, world, rank, local_rank):
        self.torch, self.dist, self.world, self.rank, self.local_rank = torch, dist, world, rank, local_rank

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX if self.world > 1 else None)

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM if self.world > 1 else None)


def resident_roofline(n, info, res_ms, prof, prof_ms, apps, nsteps, ngp, name="elastic30"):
    """roofline of k_cg_resident: the WHOLE DPCG solve of an all-elastic RVE inside one thread-block cluster (operator,
    dot products, vector updates; p and du in shared memory, r in registers).  Its bound is the FP64 pipe: a DPCG
    iteration moves no HBM byte, a solve reads b and writes du (48 B per node).  `achieved` counts the ALGORITHMIC
    flops of the operator (243 FMA per interior node and iteration) -- the kernel executes 153 of them for
    mirror-symmetric row blocks, and about 42 more FMA per node for the dot products and vector updates."""
    peak, peak_src = hbm_peak()
    sec = max(res_ms, 1e-9) * 1e-3
    nint, nn = (n - 2) ** 3, n ** 3
    flops = 2.0 * 243.0 * nint * apps
    tf = flops / sec / 1e12
    solves = ngp * nsteps
    hbm_bytes = 48.0 * nn * solves
    traffic, traffic_src = None, None
    tr = ROOT / "profiles" / "spmv_traffic.json"
    if tr.exists():   # DRAM counters need ncu's kernel replay: the per-solve byte count of the committed capture, scaled
        try:
            ent = json.loads(tr.read_text())[name]["resident"]
            traffic = ent["dram_bytes_per_rve_solve"] * ngp      # one launch = one solve of every RVE of the wave
            traffic_src = "from profiles/: " + ent["source"] + " (dram__bytes_read.sum + dram__bytes_write.sum per " \
                          "RVE solve, not measured live)"
        except Exception:
            pass
    r_extra = {"traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": 48.0 * nn * ngp}
    return {**r_extra, "kernel": "k_cg_resident (whole DPCG solve: operator + dot products + vector updates of every iteration in "
                      "ONE launch, one thread-block cluster per RVE; no HBM traffic inside the loop)",
            "bound": "fp64", "bound_detail": "FP64 DFMA pipe (vector units; BASELINE north_star: no tensor cores, FP64 "
                                             "throughout); the HBM view of the same kernel is under `hbm`",
            "achieved": tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP64_PEAK_TFLOPS,
            "peak_source": "nominal FP64: 148 SM x 64 DFMA/clk x 1965 MHz; a cluster of %d CTAs x %d clusters in flight "
                           "occupies %d of the 148 SMs" % (info["cs"], info["clusters"], info["cs"] * info["clusters"]),
            "flops_counted": "2 x 243 x interior nodes per DPCG iteration (the dense 27 x 3 x 3 stencil of SURVEY 8d); "
                             "the kernel runs the 153 structurally non-zero terms of a mirror-symmetric row block",
            "rve_applications": apps, "launches": solves / max(ngp, 1), "kernel_ms": res_ms,
            "share_of_step": res_ms / max(prof_ms, 1e-9), "instrumented_step_ms": prof_ms / nsteps,
            "us_per_rve_iteration": res_ms * 1e3 / max(apps, 1.0),
            "cluster": info,
            "hbm": {"algorithmic_bytes_per_solve": 48.0 * nn, "achieved_gbs": hbm_bytes / sec / 1e9, "peak_gbs": peak,
                    "peak_source": peak_src,
                    "three_kernel_loop_bytes_per_rve_iteration": 248.0 * nn,
                    "note": "the three-kernel loop (MICROPP_RESIDENT=0) streams 248 B per node and iteration; resident: 0"},
            "other_kernels_ms": {"asm_mat": prof["asm_mat_ms"], "asm_rhs": prof["asm_rhs_ms"],
                                 "cg_vectors": prof["cg_vec_ms"]}}


def spmv_roofline(name, n, prof, prof_ms, apps, implicit_kernel, nsteps, vec_apps=None):
    """roofline of the DPCG SpMV (+ the DPCG vector kernels) from the instrumented repeat (CUDA events per launch)"""
    peak, peak_src = hbm_peak()
    spmv_ms = prof["spmv_ms"]
    sec = max(spmv_ms, 1e-9) * 1e-3
    common = {"rve_applications": apps, "launches": prof["spmv_launches"], "kernel_ms": spmv_ms,
              "share_of_step": spmv_ms / max(prof_ms, 1e-9), "instrumented_step_ms": prof_ms / nsteps,
              "other_kernels_ms": {"asm_mat": prof["asm_mat_ms"], "asm_rhs": prof["asm_rhs_ms"],
                                   "cg_vectors": prof["cg_vec_ms"]}}
    if implicit_kernel >= 0:
        # all-elastic RVE: the Jacobian is never stored per RVE; the SpMV moves only p and Ap and is bound by the
        # FP64 pipe (243 DFMA per interior node), so BOTH fractions are reported; `frac` stays the HBM one
        b_alg = spmv_imp_bytes_per_rve(n)
        flops = 2.0 * 243.0 * (n - 2) ** 3
        achieved = b_alg * apps / sec / 1e9
        kname = {0: "k_spmv_dot_imp<8>", 3: "k_spmv_dot_tmac"}.get(implicit_kernel, f"implicit kernel {implicit_kernel}")
        key = "implicit_" + {0: "simple", 3: "tmac"}.get(implicit_kernel, str(implicit_kernel))
        r = {"kernel": kname + " (DPCG SpMV + p.Ap on the implicit elastic operator: no matrix stream)",
             "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
             "peak_source": peak_src, "traffic": None, "bytes_per_rve_application": b_alg,
             "limiter": "fp64 pipe, not HBM: the 1944 B/node matrix stream of the assembled path is eliminated, "
                        "not moved faster (see roofline_assembled for the HBM-bound SpMV of the same workload)",
             "fp64": {"achieved_tflops": flops * apps / sec / 1e12, "peak_tflops": FP64_PEAK_TFLOPS,
                      "frac": flops * apps / sec / 1e12 / FP64_PEAK_TFLOPS,
                      "peak_source": "nominal: 148 SM x 64 DFMA/clk x 1965 MHz"},
             "equivalent_664": 664.0 * 3 * n ** 3 * apps / sec / 1e9}
    else:
        b_alg, b_664 = spmv_bytes_per_rve(n)
        achieved = b_alg * apps / sec / 1e9
        key = "assembled"
        r = {"kernel": "k_spmv_dot (DPCG SpMV + p.Ap, one assembled ELL matrix per RVE)", "bound": "hbm",
             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
             "traffic": None, "achieved_664": b_664 * apps / sec / 1e9, "bytes_per_rve_application": b_alg}
    # the DPCG vector kernels (cg_update + cg_pupdate, HBM-bound streams; cg_init / cg_finish / u += du are in the
    # same timer but run once per Newton step): bytes per node and iteration from DESIGN.md section 5
    vec_b = (76.0 + 124.0 if implicit_kernel >= 0 else 120.0 + 120.0) * n ** 3
    vec_sec = max(prof["cg_vec_ms"], 1e-9) * 1e-3
    va = apps if vec_apps is None else vec_apps   # every DPCG iteration, whichever operator served it
    common["dpcg_vector_kernels"] = {"bound": "hbm", "achieved": vec_b * va / vec_sec / 1e9, "peak": peak,
                                     "unit": "GB/s", "frac": vec_b * va / vec_sec / 1e9 / peak,
                                     "bytes_per_rve_iteration": vec_b,
                                     "share_of_step": prof["cg_vec_ms"] / max(prof_ms, 1e-9)}
    r.update(common)
    # `traffic` cannot be measured inside a timed run (DRAM counters need ncu's kernel replay): it is the per-RVE DRAM
    # byte count of the committed `ncu --set full` capture of the same kernel, scaled to this run's launches
    tr = ROOT / "profiles" / "spmv_traffic.json"
    if tr.exists():
        try:
            ent = json.loads(tr.read_text())[name][key]
            r["traffic"] = ent["dram_bytes_per_rve_application"] * apps / max(prof["spmv_launches"], 1)
            r["traffic_source"] = "from profiles/: " + ent.get("source", "profiles/spmv_traffic.json") + \
                                  " (dram__bytes_read.sum + dram__bytes_write.sum per RVE application, not measured live)"
            r["algorithmic_bytes_per_launch"] = b_alg * apps / max(prof["spmv_launches"], 1)
        except Exception:
            pass
    return r


def run_b200_workload(M, env: Env, name: str, ngp: int, steps: int, warmup: int, *, cpu_baseline: bool,
                      cpu_sample, assembled_repeat: bool, full_path: bool = False):
    """One workload on this rank's GPU; returns the entry (headline line fields or a `workloads` entry)."""
    import ctypes as C
    torch = env.torch
    wl = WORKLOADS[name]
    n = wl["n"]
    rank, local_rank, world = env.rank, env.local_rank, env.world
    cfg = {"workload": wl["label"] if ngp == wl["ngp"] else wl["label"] + f" -- bounded shard of {ngp} GPs/GPU",
           "name": name, "rve_nodes": f"{n}^3", "gps_per_gpu": ngp, "coupling": "FE_ONE_WAY",
           "sharding": "independent GPs per rank, no collective",
           "l2": "inputs exceed L2: every DPCG pass streams the vectors of all resident RVEs (%.1f GB per pass at %d "
                 "GPs); assembled-matrix path: plus one %.1f MB ELL matrix per RVE"
                 % (8 * 24.0 * n ** 3 * ngp / 1e9, ngp, 1944.0 * (n - 2) ** 3 / 1e6)}
    t_ctor = time.perf_counter()
    m = M.Micropp3(M.default_params(size=(n, n, n), ngp=ngp, mpi_rank=local_rank, **wl["params"]))
    ctor_s = time.perf_counter() - t_ctor

    # pinned host buffers for the boundary crossing (the C ABI takes plain host pointers)
    eps_pin = torch.empty((ngp, 6), dtype=torch.float64).pin_memory()
    sig_pin = torch.empty((ngp, 6), dtype=torch.float64).pin_memory()
    eps_h, sig_h = eps_pin.numpy(), sig_pin.numpy()
    dp = C.POINTER(C.c_double)
    h = C.byref(m.h)
    lib = m.lib

    def e2e_step(k):
        # micropp3_set_strains / micropp3_get_stresses ARE the macro code's loop: `for gp: set_strain(gp, eps+6gp)` /
        # `for gp: get_stress(gp, sig+6gp)` compiled in C (micropp_host.cpp), i.e. the reference's per-GP calls without
        # a Python interpreter in between
        eps_h[:] = strains_for(name, ngp, rank, k)
        lib.micropp3_set_strains(h, eps_h.ctypes.data_as(dp))
        lib.micropp3_homogenize(h)
        lib.micropp3_get_stresses(h, sig_h.ctypes.data_as(dp))

    if full_path:
        # the whole load path, every step with update_vars (GPs turn non-linear at different steps)
        nsteps = wl["prep_steps"] + 4 if wl["prep_steps"] else 1
        sampler = ClockSampler(local_rank)
        env.barrier()
        sampler.start()
        l0 = m.launch_count()
        per_step, t0 = [], time.perf_counter()
        for k in range(nsteps):
            e2e_step(k)
            per_step.append({"step": k, "dev_ms": env.max(m.last_homogenize_ms()),
                             "non_linear_gps": int(env.sum(float(m.get_non_linear_gps()))),
                             "mean_cg_its": env.sum(float(np.sum([m.get_cost(g) for g in range(ngp)]))) / (ngp * world)})
            m.update_vars()
        env.barrier()
        wall = env.max(time.perf_counter() - t0)
        clocks = sampler.stop()
        launches = int(env.sum(float(m.launch_count() - l0)))
        tot = env.sum(float(ngp))
        dev = sum(p["dev_ms"] for p in per_step)
        m.close()
        return {"value": tot * nsteps / (dev * 1e-3), "unit": "GP-steps/s", "ms_per_step": dev / nsteps, "steps": nsteps,
                "config": dict(cfg, path=f"load steps 0..{nsteps - 1} with update_vars after each"),
                "e2e": {"value": tot * nsteps / wall, "unit": "GP-steps/s", "h2d_bytes_per_step": int(ngp * 48),
                        "d2h_bytes_per_step": int(ngp * 48), "ms_per_step": wall / nsteps * 1e3},
                "per_step": per_step, "gpu_launches": launches, "clocks": clocks}

    for k in range(wl["prep_steps"]):
        e2e_step(k)
        m.update_vars()
    kk = wl["prep_steps"]
    for _ in range(warmup):
        e2e_step(kk)

    # ---- timed region 1: device-event time of homogenize() with the strains already handed over ----
    sampler = ClockSampler(local_rank)
    env.barrier()
    sampler.start()
    launches0 = m.launch_count()
    dev_ms = 0.0
    for _ in range(steps):
        lib.micropp3_homogenize(h)
        dev_ms += m.last_homogenize_ms()
    env.barrier()
    launches = m.launch_count() - launches0
    cost = np.array([m.get_cost(g) for g in range(ngp)], dtype=np.float64)
    conv = sum(m.has_converged(g) for g in range(ngp))
    nl = m.get_non_linear_gps()

    # ---- timed region 2: end to end through the C ABI with host buffers ----
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step(kk)
    env.barrier()
    wall2 = time.perf_counter() - t0
    clocks = sampler.stop()
    assert np.all(np.isfinite(sig_h)), "non-finite homogenized stress"
    sig_gpu = sig_h.copy()

    # ---- instrumented repeat of the same steps: CUDA events around every kernel launch (plain stream launches) ----
    m.prof_enable(True)
    m.prof_read(True)
    prof_dev_ms = 0.0
    for _ in range(steps):
        lib.micropp3_homogenize(h)
        prof_dev_ms += m.last_homogenize_ms()
    prof = m.prof_read(True)
    res_ms = m.prof_resident_ms(True)   # cluster-resident DPCG solves (one launch per Newton step and wave)
    res_info = m.resident_info()
    m.prof_enable(False)

    dev_ms_max = env.max(dev_ms)
    wall2_max = env.max(wall2)
    total_gps = env.sum(float(ngp))
    launches_tot = int(env.sum(float(launches)))

    # roofline of the dominant kernel (rank 0's launches; every SpMV application of an RVE = one CG iteration)
    apps = float(np.sum(cost)) * steps
    imp_kernel = m.implicit_kernel()
    hybrid = None
    if imp_kernel < 0 and prof["hybrid_slot_apps"] > 0:
        # RVEs with a damage / plastic phase: slots that are mostly inside their linear regime run the HYBRID operator
        # (no matrix stream but for their listed rows).  Algorithmic bytes of one application: the implicit operator's
        # p read + Ap write for every node, plus per LISTED node its 243 values (1944 B), its list index (4 B) and the
        # read-modify-write of its Ap (48 B) -- the listed-row count is summed on the device per application.
        peak, peak_src = hbm_peak()
        h_apps, h_rows, h_ms = float(prof["hybrid_slot_apps"]), float(prof["hybrid_row_apps"]), prof["hybrid_spmv_ms"]
        h_bytes = spmv_imp_bytes_per_rve(n) * h_apps + 1996.0 * h_rows
        h_gbs = h_bytes / (max(h_ms, 1e-9) * 1e-3) / 1e9
        h_flop = 2.0 * 243.0 * ((n - 2) ** 3 * h_apps + h_rows)
        hybrid = {"kernel": "k_spmv_dot_tmac + k_spmv_fix (implicit elastic row blocks, every node) + k_spmv_hyb "
                            "(explicit rows of the nodes that touch a non-linear element)",
                  "rve_applications": h_apps, "kernel_ms": h_ms,
                  "share_of_applications": h_apps / max(apps, 1.0),
                  "share_of_step": h_ms / max(prof_dev_ms, 1e-9),
                  "mean_listed_row_fraction": h_rows / max(h_apps, 1.0) / (n - 2) ** 3,
                  "bound": "hbm", "achieved": h_gbs, "peak": peak, "unit": "GB/s", "frac": h_gbs / peak,
                  "peak_source": peak_src, "bytes_per_rve_application": h_bytes / max(h_apps, 1.0),
                  "fp64": {"achieved_tflops": h_flop / (max(h_ms, 1e-9) * 1e-3) / 1e12, "peak_tflops": FP64_PEAK_TFLOPS,
                           "frac": h_flop / (max(h_ms, 1e-9) * 1e-3) / 1e12 / FP64_PEAK_TFLOPS},
                  "assembled_equivalent_gbs": spmv_bytes_per_rve(n)[0] * h_apps / (max(h_ms, 1e-9) * 1e-3) / 1e9}
        apps = float(prof["spmv_slot_apps"])
    if res_info is not None and res_ms > 0.0:
        roof = resident_roofline(n, res_info, res_ms, prof, prof_dev_ms, apps, steps, ngp, name)
    else:
        roof = spmv_roofline(name, n, prof, prof_dev_ms, apps, imp_kernel, steps,
                             vec_apps=apps + float(prof["hybrid_slot_apps"]))
    if hybrid is not None:
        if prof["hybrid_spmv_ms"] > prof["spmv_ms"]:
            # the hybrid operator is the dominant kernel of this step: it is the roofline entry, the fully assembled
            # k_spmv_dot of the remaining slots moves under `assembled_slots`
            asm = {k: roof[k] for k in ("kernel", "achieved", "frac", "rve_applications", "launches", "kernel_ms",
                                        "share_of_step", "bytes_per_rve_application") if k in roof}
            for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "peak_source", "rve_applications",
                      "kernel_ms", "share_of_step", "bytes_per_rve_application", "fp64", "share_of_applications",
                      "mean_listed_row_fraction", "assembled_equivalent_gbs"):
                roof[k] = hybrid[k]
            roof["traffic"] = None
            roof.pop("traffic_source", None)
            roof.pop("algorithmic_bytes_per_launch", None)
            roof.pop("achieved_664", None)
            roof["limiter"] = ("mixed: the implicit part is bound by the FP64 pipe (243 DFMA per node), the listed rows "
                               "by HBM; `assembled_equivalent_gbs` is the rate the fully assembled SpMV would have "
                               "needed for the same applications in the same time")
            roof["assembled_slots"] = asm
        else:
            roof["hybrid_operator"] = hybrid

    # all-elastic workloads: the same workload once more through the assembled-matrix path (MICROPP_IMPLICIT=0) on a
    # bounded batch, instrumented, so that the HBM-bound SpMV the north star names is measured in the same run
    roof_asm = None
    if imp_kernel >= 0 and rank == 0 and assembled_repeat:
        try:
            ngp_a = min(ngp, 256)
            os.environ["MICROPP_IMPLICIT"] = "0"
            ma = M.Micropp3(M.default_params(size=(n, n, n), ngp=ngp_a, mpi_rank=local_rank, **wl["params"]))
            del os.environ["MICROPP_IMPLICIT"]
            ea = strains_for(name, ngp, rank, kk)[:ngp_a]
            ma.set_strains(ea)
            for _ in range(2):
                ma.homogenize()
            ma.prof_enable(True)
            ma.prof_read(True)
            ma.homogenize()
            ms_a = ma.last_homogenize_ms()
            prof_a = ma.prof_read(True)
            ma.prof_enable(False)
            apps_a = float(sum(ma.get_cost(g) for g in range(ngp_a)))
            roof_asm = spmv_roofline(name, n, prof_a, ms_a, apps_a, -1, 1)
            roof_asm["gps"] = ngp_a
            roof_asm["value_gps_per_s"] = ngp_a / (ms_a * 1e-3)
            sa = ma.get_stresses()
            roof_asm["max_rel_diff_vs_implicit"] = float(np.max(np.abs(sa - sig_gpu[:ngp_a]) /
                                                                np.max(np.abs(sa), axis=1, keepdims=True)))
            ma.close()
        except Exception as ex:
            os.environ.pop("MICROPP_IMPLICIT", None)
            roof_asm = {"unavailable": str(ex)}

    entry = {"value": total_gps * steps / (dev_ms_max * 1e-3), "unit": "GP/s", "steps": steps, "warmup": warmup,
             "ms_per_step": dev_ms_max / steps,
             "config": dict(cfg, newton_cg={"mean_cg_its_per_gp": float(np.mean(cost)), "converged": conv,
                                            "non_linear_gps": nl, "wave": m.wave_size()}, ctor_s=ctor_s),
             "e2e": {"value": total_gps * steps / wall2_max, "unit": "GP/s", "h2d_bytes_per_step": int(ngp * 48),
                     "d2h_bytes_per_step": int(ngp * 48), "ms_per_step": wall2_max / steps * 1e3,
                     "api": "micropp3_set_strains + micropp3_homogenize + micropp3_get_stresses on host buffers: the "
                            "first and last are C loops over the reference's per-GP set_strain / get_stress"},
             "gpu_launches": launches_tot, "clocks": clocks, "roofline": roof}
    if roof_asm is not None:
        entry["roofline_assembled"] = roof_asm

    if rank == 0 and world == 1 and cpu_baseline:
        try:
            cores = host_cores()
            restart = None
            if wl["prep_steps"]:
                # bounded sample: the state after the preparation steps of the sample's GPs is produced by the product
                # (a second small object: steps 0..prep-1 on the GPU) and handed to the reference as a restart file
                sample = cpu_sample or min(cores, 8)
                import tempfile
                tmp = tempfile.mkdtemp(prefix="micropp_bench_")
                cwd = os.getcwd()
                os.chdir(tmp)
                try:
                    ms_ = M.Micropp3(M.default_params(size=(n, n, n), ngp=sample, mpi_rank=0, **wl["params"]))
                    for k in range(wl["prep_steps"]):
                        ms_.set_strains(strains_for(name, sample, 0, k))
                        ms_.homogenize()
                        ms_.update_vars()
                    ms_.write_restart(1)
                    ms_.close()
                finally:
                    os.chdir(cwd)
                restart = (tmp, 1)
            else:
                sample = cpu_sample or 4 * cores
            res = run_reference(name, wl, sample, 1, 0, restart_from=restart)
            if restart is not None:
                import shutil
                shutil.rmtree(restart[0], ignore_errors=True)
            entry["cpu_baseline"] = {"value": res["value"], "unit": "GP/s", "cores": res["cores"], "kind": "reference",
                                     "sample": res["sample"]}
            # the sample's GPs are the first GPs of rank 0's batch: same strains => the stresses must agree
            ref_sig = res["stress"]
            d = np.max(np.abs(ref_sig - sig_gpu[:sample]), axis=1) / np.maximum(np.max(np.abs(ref_sig), axis=1), 1e-300)
            entry["cpu_baseline"]["max_rel_stress_diff_vs_gpu"] = float(np.max(d))
        except Exception as ex:  # the baseline is reported, never required for the GPU number
            entry["cpu_baseline"] = {"value": None, "unit": "GP/s", "cores": host_cores(), "kind": "reference",
                                     "sample": f"unavailable: {ex}"}
    m.close()
    return entry


def run_slab(M, env: Env, n: int, reps: int):
    """BASELINE configs[4]: ONE n^3 elastic RVE (sphere, contrast 10, eps = {1e-3,0,0,0,0,0}) over z-slabs on all ranks'
    GPUs (strong scaling): halo planes of p and the DPCG dot products travel over NVLink peer memory."""
    from micropp_b200.slab import SlabRVE
    torch = env.torch
    kw = dict(size=(n, n, n), type=1, geo_params=(0.2, 0, 0, 0), materials=[EL(1e7), EL(1e8), EL(1e7)],
              lin_stress=False, calc_ctan_lin=False)
    eps = np.array([1e-3, 0, 0, 0, 0, 0.0])
    w = (env.dist, env.rank, env.world) if env.world > 1 else None
    t0 = time.perf_counter()
    rve = SlabRVE(kw, world=w, nslabs=1, device=env.local_rank)
    ctor = time.perf_counter() - t0
    times = []
    out = None
    l0 = rve.launch_count()
    for _ in range(reps + 1):          # the first solve also builds the CUDA graphs: untimed
        env.barrier()
        t0 = time.perf_counter()
        out = rve.homogenize(eps)
        torch.cuda.synchronize()
        times.append(env.max(time.perf_counter() - t0))
    launches = int(env.sum(float(rve.launch_count() - l0)))
    best = min(times[1:])
    its = out["cg_its"]
    entry = {"value": best * 1e3, "unit": "ms per homogenize() of the one RVE", "higher_is_better": False,
             "scaling": "strong", "n_gpus": env.world, "ms_all": [round(x * 1e3, 2) for x in times],
             "cg_its": its, "newton_its": out["newton_its"], "converged": out["converged"],
             "stress": [float(x) for x in out["stress"]], "cg_iteration_us": best * 1e6 / max(its, 1),
             "config": {"workload": f"configs[4]: single {n}^3-node RVE, elastic sphere contrast 10, z-slabs over "
                                    f"{env.world} GPU(s), implicit operator", "exchange": rve.exchange,
                        "halo_bytes_each_way_per_iteration": 24 * n * n, "ctor_s": ctor},
             "peer_error": rve.peer_error(), "gpu_launches": launches}
    fx = ROOT / "tests" / "golden" / f"bench_slab{n}.npz"
    if fx.exists():   # the reference's own result for this RVE (tests/golden/make_golden_full.py)
        f = np.load(fx)
        entry["vs_reference_fixture"] = {"stress_relerr": float(np.max(np.abs(out["stress"] - f["sig"])) /
                                                                np.max(np.abs(f["sig"]))),
                                         "reference_cg_its": int(f["cost"])}
    rve.close()
    return entry


# ------------------------------------------------------------------------------------------------ main
METRIC = "homogenized GPs/sec (DPCG+assembly)"
EXTRA_NGP = {"damage50": 128, "plastic40": 128, "elastic30": 1024}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["slab200"], default="elastic30")
    ap.add_argument("--ngp", type=int, default=None, help="GPs per GPU (default: the workload's)")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--cpu-sample", type=int, default=None, help="GPs in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-assembled", action="store_true",
                    help="skip the assembled-matrix repeat of an all-elastic workload (roofline_assembled)")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (no `workloads` entries)")
    ap.add_argument("--extra", default="damage50,plastic40,slab200", help="comma list of the additional workloads")
    ap.add_argument("--full-path", action="store_true", help="time the whole load path of the workload")
    ap.add_argument("--slab-n", type=int, default=200)
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and args.impl == "b200":
        # not under torchrun: launch ourselves the way the driver would
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               str(Path(__file__).resolve())] + sys.argv[1:]
        os.execv(sys.executable, cmd)

    warmup = max(args.warmup, 0)

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        name = args.workload if args.workload in WORKLOADS else "elastic30"
        wl = WORKLOADS[name]
        steps = args.steps if args.steps is not None else (20 if name == "elastic30" else 3)
        n = wl["n"]
        ngp = args.ngp or wl["ngp"]
        cores = host_cores()
        sample = args.cpu_sample or (4 * cores if name == "elastic30" else cores)
        res = run_reference(name, wl, sample, steps, warmup)
        base_cfg = {"workload": wl["label"], "name": name, "rve_nodes": f"{n}^3", "gps_per_gpu": ngp,
                    "coupling": "FE_ONE_WAY", "sharding": "independent GPs per rank, no collective",
                    "l2": "inputs exceed L2: every DPCG pass streams the vectors of all resident RVEs (%.1f GB per pass "
                          "at %d GPs); assembled-matrix path: plus one %.1f MB ELL matrix per RVE"
                          % (8 * 24.0 * n ** 3 * ngp / 1e9, ngp, 1944.0 * (n - 2) ** 3 / 1e6)}
        line = {"impl": "reference", "metric": METRIC, "value": res["value"],
                "unit": "GP/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": base_cfg,
                "cpu_baseline": {"value": res["value"], "unit": "GP/s", "cores": res["cores"], "kind": "reference",
                                 "sample": res["sample"]},
                "e2e": {"value": res["value"], "unit": "GP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = Env(torch, dist, world, rank, local_rank)

    import micropp_b200 as M
    M.load()

    if args.workload == "slab200":
        e = run_slab(M, env, args.slab_n, max(args.steps or 2, 1))
        line = {"metric": "ms per homogenize() of one 200^3 RVE over z-slabs", "warmup": 1, "steps": len(e["ms_all"]) - 1,
                "ms_per_step": e["value"], "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        line.update(e)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return 0

    name = args.workload
    wl = WORKLOADS[name]
    steps = args.steps if args.steps is not None else (20 if name == "elastic30" else 3)
    ngp = args.ngp or wl["ngp"]
    entry = run_b200_workload(M, env, name, ngp, steps, warmup, cpu_baseline=not args.no_cpu_baseline,
                              cpu_sample=args.cpu_sample, assembled_repeat=not args.no_assembled,
                              full_path=args.full_path)
    line = {"metric": METRIC, "value": entry.pop("value"), "unit": entry.pop("unit"), "n_gpus": world,
            "steps": entry.pop("steps"), "warmup": warmup, "ms_per_step": entry.pop("ms_per_step"),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    line.update(entry)

    # ---------------------------------------------------------------- the other BASELINE configs, bounded
    if not args.no_extra and not args.full_path:
        extras = {}
        for x in [t for t in args.extra.split(",") if t and t != name]:
            t0 = time.perf_counter()
            try:
                if x == "slab200":
                    extras[x] = run_slab(M, env, args.slab_n, 2)
                elif x in WORKLOADS:
                    extras[x] = run_b200_workload(M, env, x, EXTRA_NGP[x], 2, 3, cpu_baseline=not args.no_cpu_baseline,
                                                  cpu_sample=None, assembled_repeat=False)
                    extras[x].update(metric=METRIC, higher_is_better=True, scaling="weak", n_gpus=world)
            except Exception as ex:   # an extra workload never takes the headline down
                extras[x] = {"unavailable": f"{type(ex).__name__}: {ex}"}
            extras[x]["bench_wall_s"] = time.perf_counter() - t0
        line["workloads"] = extras

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
