"""GPU diagnostic of the cluster-resident DPCG kernel: one small batch per shape / cluster size against the three-kernel
loop (MICROPP_RESIDENT=0); prints iteration counts and stress differences.  python tools/resident_check.py"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import micropp_b200 as M  # noqa: E402
from common import CASES, relerr  # noqa: E402


def run(dims, ngp, eps, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        g = M.Micropp3(M.default_params(size=dims, ngp=ngp, lin_stress=False, calc_ctan_lin=False,
                                        **CASES["elastic_sphere"]))
        for gp in range(ngp):
            g.set_strain(gp, eps[gp])
        g.homogenize()
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    sig = np.array([g.get_stress(gp) for gp in range(ngp)])
    return g, sig, [g.get_cost(gp) for gp in range(ngp)], g.last_homogenize_ms()


for dims, ngp in [((6, 6, 6), 2), ((12, 12, 12), 3), ((14, 12, 18), 3), ((20, 20, 20), 4), ((30, 30, 30), 64)]:
    eps = np.random.default_rng(1).uniform(-1e-3, 1e-3, (ngp, 6))
    gl, sl, cl, tl = run(dims, ngp, eps, {"MICROPP_RESIDENT": "0"})
    for cs in ("0", "1", "2", "4", "8"):
        env = {"MICROPP_VERBOSE": "1"}
        if cs != "0":
            env["MICROPP_RESIDENT_CS"] = cs
        g, s, c, t = run(dims, ngp, eps, env)
        info = g.resident_info()
        if info is None:
            print(dims, "cs", cs, "no plan", flush=True)
            continue
        print(dims, "cs", cs, info, "its", c[:4], "loop its", cl[:4], "relerr %.2e" % max(relerr(s[i], sl[i]) for i in
              range(ngp)), "ms %.3f vs loop %.3f" % (t, tl), flush=True)
        del g
