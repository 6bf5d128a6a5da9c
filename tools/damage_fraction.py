"""VERDICT r1 task 6, first half: how much of a damage50 RVE is actually damaged?

For a hybrid operator (rows of undamaged neighbourhoods served from the implicit elastic row blocks, only damaged rows as
explicit ELL tiles) to pay off, the fraction of interior rows whose 8 elements hold ANY Gauss point past the damage threshold (D != 0) must be
well below 1 for the Gauss points that cost the time.  This script runs bench.py's damage50 load path on a few Gauss
points at 50^3, reads the internal variables back after every load step (reference layout [elem][gp][7]: r, D, ...) and
prints, per load step and Gauss point: that fraction, and the CG iterations the step cost (the weight).

    python tools/damage_fraction.py [ngp=12] [steps=10]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import micropp_b200 as M

ngp = int(sys.argv[1]) if len(sys.argv) > 1 else 12
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
wl = bench.WORKLOADS["damage50"]
n = wl["n"]
M.load()
m = M.Micropp3(M.default_params(size=(n, n, n), ngp=ngp, **wl["params"]))
ne = n - 1
out = []
for k in range(steps):
    m.set_strains(bench.strains_for("damage50", wl["ngp"], 0, k)[:ngp])
    m.homogenize()
    cost = [m.get_cost(g) for g in range(ngp)]
    m.update_vars()
    frac = []
    for g in range(ngp):
        v = m.get_vars(g, 0)
        if v is None:
            frac.append(0.0)
            continue
        D = v.reshape(ne, ne, ne, 8, 7)[..., 1]                 # [ez][ey][ex][gp]
        dam = (D != 0).any(axis=3)                              # element has a Gauss point past the threshold (D = 1 - q/r != 0;
                                                                # with the hard-coded H0 = 10 the law hardens: D < 0)
        # interior node (i, j, k) touches elements (i-1..i, j-1..j, k-1..k)
        node = np.zeros((n - 2, n - 2, n - 2), dtype=bool)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    node |= dam[dz:dz + n - 2, dy:dy + n - 2, dx:dx + n - 2]
        frac.append(float(node.mean()))
    w = np.array(cost, dtype=float)
    row = dict(step=k, mean_fraction=float(np.mean(frac)), cost_weighted_fraction=float(np.dot(frac, w) / max(w.sum(), 1)),
               fractions=[round(f, 3) for f in frac], cost=cost)
    out.append(row)
    print(json.dumps(row), flush=True)
print(json.dumps({"workload": "damage50", "ngp": ngp, "per_step": out}))
