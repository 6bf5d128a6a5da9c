"""GPU: per-warp phase timeline of the cluster-resident DPCG kernel (dbg bit 256: clock64 counters per warp and phase,
60 forced iterations).  Prints, per CTA of the cluster that solved slot 0, the cycles per iteration of each phase for
the slowest / fastest warp.  python tools/resident_timeline.py [n] [ngp]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import micropp_b200 as M  # noqa: E402
from common import CASES  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ngp = int(sys.argv[2]) if len(sys.argv) > 2 else 15
g = M.Micropp3(M.default_params(size=(n, n, n), ngp=ngp, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"]))
eps = np.random.default_rng(1).uniform(-1e-3, 1e-3, (ngp, 6))
for gp in range(ngp):
    g.set_strain(gp, eps[gp])
g.homogenize()
info = g.resident_info()
ms = g.bench_resident(ngp, 1, 4 | 256)
t = g.resident_timeline(0).astype(np.float64) / 60.0
names = ["interface", "operator", "wait blk", "barrier A", "update", "barrier B", "du/p+push", "barrier C"]
print(info, "ms", ms, "-> %.2f us per iteration" % (ms * 1e3 / 60))
nw = info["threads"] // 32
print("cycles per iteration; rows: CTA rank; per phase: warp values (w0..w%d)" % (nw - 1))
for r in range(info["cs"]):
    print("CTA", r, "total per warp", np.round(t[r, :nw].sum(axis=1)).astype(int).tolist())
    for k, nm in enumerate(names):
        print("    %-10s" % nm, np.round(t[r, :nw, k]).astype(int).tolist())
