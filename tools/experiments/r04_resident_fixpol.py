"""GPU timing experiment: interface-group -> warp assignment policies (MICROPP_RES_FIXPOL) and interface pass before /
after the operator pass (dbg bit 128) of the cluster-resident DPCG kernel.  python tools/resident_fixpol.py"""
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
for dbg in (4, 4 | 128, 5):
    ms = g.bench_resident(ngp, 3, dbg)
    print("   dbg %3d: %.2f us per iteration and wave" % (dbg, ms * 1e3 / 60 / 4), flush=True)
'''
for pol in ("0", "1", "2", "3"):
    print("MICROPP_RES_FIXPOL =", pol, flush=True)
    env = dict(os.environ, MICROPP_RES_FIXPOL=pol)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("\n".join(l for l in out.stdout.splitlines() if "dbg" in l), out.stderr[-300:] if out.returncode else "", flush=True)
