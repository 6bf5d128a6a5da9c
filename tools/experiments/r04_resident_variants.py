"""GPU timing experiment: library variants (MICROPP_B200_LIB) of the cluster-resident DPCG kernel, isolated launches with
exactly 60 iterations.  python tools/resident_variants.py lib1.so lib2.so ..."""
import os
import subprocess
import sys

code = r'''
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import micropp_b200 as M
from common import CASES
n, ngp = 30, 60
g = M.Micropp3(M.default_params(size=(n, n, n), ngp=ngp, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"]))
eps = np.random.default_rng(1).uniform(-1e-3, 1e-3, (ngp, 6))
for gp in range(ngp): g.set_strain(gp, eps[gp])
g.homogenize()
print("   its", g.get_cost(0), "stress", g.get_stress(0)[:2])
for dbg in (4, 5, 4, 5):
    ms = g.bench_resident(ngp, 3, dbg)
    print("   dbg %3d: %.2f us per iteration and wave" % (dbg, ms * 1e3 / 60 / 4), flush=True)
'''
for lib in sys.argv[1:]:
    print(lib, flush=True)
    env = dict(os.environ, MICROPP_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("\n".join(l for l in out.stdout.splitlines() if l.startswith("   ")), out.stderr[-300:] if out.returncode else "", flush=True)
