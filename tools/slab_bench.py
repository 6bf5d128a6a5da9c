"""BASELINE configs[4]: ONE n^3 RVE (elastic sphere, contrast 10, eps = {1e-3,0,0,0,0,0}) solved over z-slabs on
N GPUs (one process per GPU; NCCL halo exchange of p + all-reduce of the DPCG dot products).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_bench.py --rve 200
    python tools/slab_bench.py --rve 50 --check     # N=1, plus parity against the single-domain batched solver

Strong scaling: the RVE is fixed, the slab per GPU shrinks with N.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rve", type=int, default=200)
    ap.add_argument("--check", action="store_true", help="compare with the single-domain solver (n <= 60)")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                    help="peer: NVLink P2P pulls + mailbox sums with device-side flags; nccl: send/recv + all-reduce")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    w = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        w = (dist, rank, world)
    import micropp_b200 as M
    from micropp_b200.slab import SlabRVE
    n = args.rve
    el = lambda E: (0, E, 0.3, 0.0, 0.0, 0.0)
    kw = dict(size=(n, n, n), type=1, geo_params=(0.2, 0, 0, 0), materials=[el(1e7), el(1e8), el(1e7)],
              lin_stress=False, calc_ctan_lin=False)
    eps = np.array([1e-3, 0, 0, 0, 0, 0.0])
    t0 = time.perf_counter()
    rve = SlabRVE(kw, world=w, nslabs=1, device=lr, exchange=args.exchange)
    ctor = time.perf_counter() - t0
    times = []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = rve.homogenize(eps)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    best = min(times)
    its = out["cg_its"]
    line = {"config": f"single {n}^3 RVE, elastic sphere contrast 10, z-slabs", "n_gpus": world,
            "exchange": args.exchange, "peer_error": rve.peer_error(), "ms": best * 1e3,
            "ms_all": [round(x * 1e3, 2) for x in times], "cg_its": its, "newton_its": out["newton_its"],
            "converged": out["converged"], "stress": [float(x) for x in out["stress"]],
            "cg_iteration_us": best * 1e6 / max(its, 1),
            "spmv_algorithmic_GBps_total": 1992.0 * (n - 2) ** 3 * its / best / 1e9,
            "halo_exchanges": rve.exchanges // args.reps, "allreduces": rve.allreduces // args.reps,
            "halo_bytes_each_way": 24 * n * n, "ctor_s": ctor}
    if args.check and rank == 0:
        one = M.Micropp3(M.default_params(**kw))
        one.set_strain(0, eps)
        one.homogenize()
        s1 = one.get_stress(0)
        line["check"] = {"single_domain_cost": one.get_cost(0),
                         "stress_relerr": float(np.max(np.abs(out["stress"] - s1)) / np.max(np.abs(s1)))}
        one.close()
    rve.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
