"""Isolated timing of the implicit elastic SpMV kernels on one B200.

    python tools/bench_imp_spmv.py [n=30] [ngp=1024] [iters=20] [case=elastic_sphere|homog]

Prints ms per application of the operator to `ngp` RVEs (SpMV + its p.Ap fold), FP64 TFLOP/s counted with the
ALGORITHMIC 2*243 flop per interior node, and the fraction of the nominal FP64 peak (148 SM x 64 DFMA/clk x 1965 MHz
= 37.2 TFLOP/s), for k_spmv_dot_imp (0) and k_spmv_dot_tmac + k_spmv_fix (3), after comparing them on a random vector
(same bits).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import micropp_b200 as M
from common import CASES

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ngp = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
case = sys.argv[4] if len(sys.argv) > 4 else "elastic_sphere"   # or "homog": one material, no interface nodes
CASES["homog"] = dict(type=0, materials=[(0, 1e7, 0.3, 0.0, 0.0, 0.0)] * 3)
M.load()
m = M.Micropp3(M.default_params(size=(n, n, n), ngp=ngp, lin_stress=False, calc_ctan_lin=False, **CASES[case]))
rng = np.random.default_rng(1)
m.set_strains(rng.uniform(-1e-3, 1e-3, (ngp, 6)))
m.homogenize()   # leaves a real search direction p in every slot
p = rng.uniform(-1, 1, m.nndim)
y0, d0 = m.apply_operator(p, op=3, kernel=0)
flop = 2.0 * 243 * (n - 2) ** 3 * ngp
peak = 148 * 64 * 2 * 1.965e9
names = {0: "k_spmv_dot_imp (table-driven, 8 slots per thread)",
         3: "k_spmv_dot_tmac + k_spmv_fix (TMA load + TMA store per tile; interface nodes by the table-driven kernel)"}
from bench import ClockSampler
only = [int(x) for x in os.environ.get("KERNELS", "0,3").split(",")]
long_iters = int(os.environ.get("LONG_ITERS", "0"))   # > 0: one long run per kernel with nvidia-smi clock sampling
out = []
for v in only:
    y, d = m.apply_operator(p, op=3, kernel=v)
    err = float(np.max(np.abs(y - y0)) / np.max(np.abs(y0)))
    ms = min(m.bench_imp_spmv(ngp, iters, v) for _ in range(3))
    clk = None
    if long_iters:
        cs = ClockSampler(0)
        cs.start()
        ms = m.bench_imp_spmv(ngp, long_iters, v)
        clk = cs.stop()
    tf = flop / (ms * 1e-3) / 1e12
    out.append(dict(kernel=v, name=names[v], ms=ms, tflops=tf, frac_fp64_peak=tf * 1e12 / peak, relerr_vs_table=err,
                    dot_relerr=abs(d - d0) / abs(d0), clocks=clk))
    print(f"kernel {v}: {ms:8.3f} ms  {tf:6.2f} TFLOP/s  {tf*1e12/peak:5.1%} of FP64 peak  err={err:.1e}  {names[v]}  "
          f"clocks={clk}", flush=True)
print(json.dumps({"case": case, "implicit_kernel": m.implicit_kernel(),
                  "rve": n, "ngp": ngp, "iters": iters, "kernels": out}))
