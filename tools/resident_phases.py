"""GPU timing experiment: the cluster-resident DPCG kernel launched alone (mgpu_bench_resident) with phases switched off
one by one (dbg bits: 1 no interface pass, 2 no halo pushes, 4 exactly 60 iterations, 8 no operator pass, 16 no du / p
update; 256 = per-warp phase counters, tools/resident_timeline.py).  Results are meaningless with any bit set -- only the time is read.
python tools/resident_phases.py [n] [ngp]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import micropp_b200 as M  # noqa: E402
from common import CASES  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ngp = int(sys.argv[2]) if len(sys.argv) > 2 else 60
eps = np.random.default_rng(1).uniform(-1e-3, 1e-3, (ngp, 6))
g = M.Micropp3(M.default_params(size=(n, n, n), ngp=ngp, lin_stress=False, calc_ctan_lin=False, **CASES["elastic_sphere"]))
for gp in range(ngp):
    g.set_strain(gp, eps[gp])
g.homogenize()
info = g.resident_info()
print(info, "homogenize ms", g.last_homogenize_ms(), "its", g.get_cost(0))
waves = -(-ngp // info["clusters"])
names = {1: "no interface pass", 2: "no pushes", 8: "no operator", 16: "no du/p update"}
for dbg in [4, 4 | 1, 4 | 2, 4 | 1 | 2, 4 | 8, 4 | 8 | 1, 4 | 8 | 1 | 2, 4 | 8 | 1 | 2 | 16, 4 | 16, 4 | 1 | 16]:
    ms = g.bench_resident(ngp, 3, dbg)
    what = ", ".join(v for k, v in names.items() if dbg & k) or "everything"
    print(f"dbg {dbg:3d} ms {ms:8.3f} -> {ms * 1e3 / 60 / waves:7.2f} us per iteration and wave of {info['clusters']} "
          f"clusters ({waves} waves)  [{what}]", flush=True)
