#!/bin/bash
# Round-end record on one B200: default bench (both arms), the other workloads, launch lists.
tag=${1:-final}
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py > $O/${tag}_bench_elastic30.json 2> $O/${tag}_bench_elastic30.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/${tag}_bench_ref_elastic30.json 2> $O/${tag}_bench_ref.err
timeout 1200 python bench.py --workload damage50 --steps 2 --warmup 3 --cpu-sample 16 > $O/${tag}_bench_damage50_ngp512.json 2> $O/${tag}_bench_damage50.err
timeout 900 python bench.py --workload plastic40 --steps 2 --warmup 3 --cpu-sample 16 > $O/${tag}_bench_plastic40_ngp256.json 2> $O/${tag}_bench_plastic40.err
MICROPP_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file $O/${tag}_launches_damage50_ngp8.csv \
    python bench.py --workload damage50 --ngp 8 --steps 1 --warmup 1 --no-cpu-baseline > $O/${tag}_ncu_launch_dmg.log 2>&1
for f in $O/${tag}_bench_*.json; do echo "== $f"; tail -c 400 $f; echo; done
