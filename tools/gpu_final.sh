#!/bin/bash
# Round-end record on one B200: default bench (both arms), the other workloads, launch list + ncu of the final kernels.
tag=${1:-final}
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py > $O/${tag}_bench_elastic30.json 2> $O/${tag}_bench_elastic30.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/${tag}_bench_ref_elastic30.json 2> $O/${tag}_bench_ref.err
timeout 900 python bench.py --workload damage50 --steps 2 --warmup 2 --no-cpu-baseline > $O/${tag}_bench_damage50_ngp512.json 2> $O/${tag}_bench_damage50.err
timeout 600 python bench.py --workload plastic40 --steps 2 --warmup 2 --no-cpu-baseline > $O/${tag}_bench_plastic40_ngp256.json 2> $O/${tag}_bench_plastic40.err
MICROPP_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file $O/${tag}_launches_elastic30_ngp64.csv \
    python bench.py --ngp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-assembled > $O/${tag}_ncu_launch.log 2>&1
MICROPP_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"k_spmv_dot_tmac|k_cg_update_imp|k_cg_pupdate_imp" -s 90 -c 3 -f -o $O/${tag}_ncu_dpcg_elastic30_ngp1024 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-assembled > $O/${tag}_ncu_full.log 2>&1
for f in $O/${tag}_bench_*.json; do echo "== $f"; tail -c 300 $f; echo; done
