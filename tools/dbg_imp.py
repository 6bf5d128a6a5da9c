import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import micropp_b200 as M
from common import CASES
M.load()
dims = (14, 9, 8)
def run(ngp, env, seed=43):
    rng = np.random.default_rng(seed)
    eps = rng.uniform(-1e-3, 1e-3, (3, 6))[:ngp]
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    g = M.Micropp3(M.default_params(size=dims, ngp=ngp, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"]))
    for k, v in old.items():
        if v is None: del os.environ[k]
        else: os.environ[k] = v
    for gp in range(ngp): g.set_strain(gp, eps[gp])
    g.homogenize()
    s = np.array([g.get_stress(gp) for gp in range(ngp)])
    c = [g.get_cost(gp) for gp in range(ngp)]
    g.close()
    return s, c
for ngp in (1, 3):
    for graphs in ("1", "0"):
        ref, cr = run(ngp, {"MICROPP_IMPLICIT": "0", "MICROPP_GRAPHS": graphs})
        for kern in ("0", "1"):
            s, c = run(ngp, {"MICROPP_IMPLICIT": "1", "MICROPP_IMP_KERNEL": kern, "MICROPP_GRAPHS": graphs})
            err = np.max(np.abs(s - ref), axis=1) / np.max(np.abs(ref), axis=1)
            print(f"ngp={ngp} graphs={graphs} kernel={kern} cost={c} ref_cost={cr} relerr={err}")
from oracle import refpy
rng = np.random.default_rng(43)
eps = rng.uniform(-1e-3, 1e-3, (3, 6))
r = refpy.RefMicropp(refpy.default_params(size=dims, ngp=3, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"]))
for gp in range(3): r.set_strain(gp, eps[gp])
r.homogenize()
sr = np.array([r.get_stress(gp) for gp in range(3)])
print("ref cost", [r.get_cost(gp) for gp in range(3)])
for name, env in (("explicit", {"MICROPP_IMPLICIT": "0"}), ("simple", {"MICROPP_IMPLICIT": "1", "MICROPP_IMP_KERNEL": "0"}), ("tiled", {"MICROPP_IMPLICIT": "1", "MICROPP_IMP_KERNEL": "1"})):
    s, c = run(3, env)
    print(name, "vs reference:", np.max(np.abs(s - sr), axis=1) / np.max(np.abs(sr), axis=1))
