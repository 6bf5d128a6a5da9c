"""Isolated timing of the implicit elastic SpMV variants (k_spmv_dot_tma / k_spmv_dot_tmac) on one B200.

    python tools/bench_imp_spmv.py [n=30] [ngp=1024] [iters=20]

Prints ms per application of the operator to `ngp` RVEs, FP64 TFLOP/s (2*243 flop per interior node) and the
fraction of the nominal FP64 peak (148 SM x 64 DFMA/clk x 1965 MHz = 37.2 TFLOP/s) for every variant, after checking
that every variant returns the bits of the simple kernel on a random vector.
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
names = {0: "k_spmv_dot_tma (rows in smem, 8 nodes/thread, 2 stages, 2 blocks/SM)",
         1: "k_spmv_dot_tmac 2 stages, 2 blocks/SM, unroll 1", 2: "k_spmv_dot_tmac 1 stage, 3 blocks/SM, unroll 1",
         3: "k_spmv_dot_tmac 1 stage, 4 blocks/SM, unroll 1", 4: "k_spmv_dot_tmac 2 stages, 2 blocks/SM, unroll 3"}
from bench import ClockSampler
only = [int(x) for x in os.environ.get("VARIANTS", "0,1,2,3,4").split(",")]
long_iters = int(os.environ.get("LONG_ITERS", "0"))   # > 0: one long run per variant with nvidia-smi clock sampling
out = []
for v in only:
    y, d = m.apply_operator(p, op=3, kernel=10 + v)
    same = bool(np.array_equal(y, y0))
    ms = min(m.bench_imp_spmv(ngp, iters, 10 + v) for _ in range(3))
    clk = None
    if long_iters:
        cs = ClockSampler(0)
        cs.start()
        ms = m.bench_imp_spmv(ngp, long_iters, 10 + v)
        clk = cs.stop()
    tf = flop / (ms * 1e-3) / 1e12
    if v >= 1 and os.environ.get("SPLIT"):
        t_nocompute = min(m.bench_imp_spmv(ngp, iters, 110 + v) for _ in range(2))
        t_noload = min(m.bench_imp_spmv(ngp, iters, 210 + v) for _ in range(2))
        print(f"   variant {v}: loads + stores only {t_nocompute:.3f} ms; compute on a stale brick only {t_noload:.3f} ms", flush=True)
    out.append(dict(variant=v, name=names[v], ms=ms, tflops=tf, frac_fp64_peak=tf * 1e12 / peak, bit_identical=same, clocks=clk))
    print(f"variant {v}: {ms:8.3f} ms  {tf:6.2f} TFLOP/s  {tf*1e12/peak:5.1%} of FP64 peak  bits_ok={same}  {names[v]}  clocks={clk}", flush=True)
print(json.dumps({"case": case, "fix_nodes": m.lib.micropp3x_implicit_rows(__import__("ctypes").byref(m.h)), "rve": n, "ngp": ngp, "iters": iters, "variants": out}))
