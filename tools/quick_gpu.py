"""Quick GPU probe: elastic sphere batch (BASELINE config 2 shape), timing breakdown by kernel class."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import micropp_b200 as M
sys.path.insert(0, "tests")
from common import CASES

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ngp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
case = sys.argv[3] if len(sys.argv) > 3 else "elastic_sphere"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2

t0 = time.time()
m = M.Micropp3(M.default_params(size=(n, n, n), ngp=ngp, lin_stress=False, calc_ctan_lin=False, nr_max_its=12,
                                **CASES[case]))
print(f"ctor {time.time()-t0:.3f}s wave={m.wave_size()}", flush=True)
rng = np.random.default_rng(1234)
if case == "elastic_sphere":
    eps = rng.uniform(-1e-3, 1e-3, (ngp, 6))
else:
    eps = np.zeros((ngp, 6))
    eps[:, 0] = rng.uniform(0.5, 1.5, ngp) * 0.012
m.prof_enable(True)
for s in range(steps):
    m.set_strains(eps * (s + 1) if case != "elastic_sphere" else eps)
    t0 = time.time()
    m.homogenize()
    wall = time.time() - t0
    pr = m.prof_read(True)
    cost = [m.get_cost(g) for g in range(ngp)]
    rows = 3 * n ** 3
    spmv_bytes = 664.0 * rows * pr["spmv_slot_apps"]
    print(f"step {s}: wall {wall*1e3:.1f} ms, dev {m.last_homogenize_ms():.1f} ms, GP/s {ngp/wall:.1f}, "
          f"cost mean {np.mean(cost):.1f} [{min(cost)},{max(cost)}], conv {sum(m.has_converged(g) for g in range(ngp))}/{ngp}, "
          f"NL {m.get_non_linear_gps()}")
    print(f"   spmv {pr['spmv_ms']:.1f} ms in {pr['spmv_launches']} launches, slot-apps {pr['spmv_slot_apps']} "
          f"=> {spmv_bytes/max(pr['spmv_ms'],1e-9)/1e6:.0f} GB/s algorithmic; asm_mat {pr['asm_mat_ms']:.1f} ms, "
          f"asm_rhs {pr['asm_rhs_ms']:.1f} ms, cg_vec {pr['cg_vec_ms']:.1f} ms", flush=True)
    m.update_vars()
m.prof_enable(False)
m.set_strains(eps)
for rep in range(3):
    t0 = time.time(); m.homogenize(); wall = time.time() - t0
    print(f"no-prof[{rep}]: wall {wall*1e3:.2f} ms  GP/s {ngp/wall:.1f}  launches so far {m.launch_count()}")
nb = min(ngp, m.wave_size())
if m.implicit_kernel() >= 0:   # all-elastic RVE: there is no per-slot matrix; the SpMV is the implicit operator's
    ms = m.bench_imp_spmv(nb, 10, 2)
    print(f"isolated implicit spmv: {ms:.3f} ms for {nb} slots => {486.0*(n-2)**3*nb/ms/1e9:.2f} TFLOP/s")
else:
    ms = m.bench_spmv(nb, 10)
    print(f"isolated spmv: {ms:.3f} ms for {nb} slots => {664.0*3*n**3*nb/ms/1e6:.0f} GB/s algorithmic")
