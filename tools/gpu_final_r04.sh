#!/bin/bash
# Round-end record on one B200 (round 2, resident DPCG): smoke, default bench, launch list + ncu of the resident kernel
tag=${1:-r04w}
O=gpurun_out
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${tag}_smoke.log
timeout 900 python bench.py > $O/${tag}_bench_default.json 2> $O/${tag}_bench_default.err
MICROPP_RESIDENT=0 timeout 300 python bench.py --no-extra --no-cpu-baseline --no-assembled --steps 10 > $O/${tag}_bench_elastic30_three_kernel_loop.json 2> $O/${tag}_bench_loop.err
MICROPP_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file $O/${tag}_launches_elastic30_ngp64.csv \
    python bench.py --ngp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-assembled --no-extra > $O/${tag}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_resident -s 1 -c 1 -f \
    -o $O/${tag}_ncu_resident_elastic30_ngp1024 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-assembled --no-extra > $O/${tag}_ncu_full.log 2>&1
ncu -i $O/${tag}_ncu_resident_elastic30_ngp1024.ncu-rep --page details > $O/${tag}_ncu_resident_elastic30_ngp1024_details.txt 2>&1
tail -3 $O/${tag}_smoke.log
python - <<PY
import json
for f in ("default", "elastic30_three_kernel_loop"):
    try:
        d = json.loads(open("$O/${tag}_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"), d["clocks"])
        r = d.get("roofline", {})
        print("   roofline", {k: r.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "share_of_step", "us_per_rve_iteration")})
        for k, v in d.get("workloads", {}).items():
            print("   ", k, v.get("value"), v.get("unit"), v.get("bench_wall_s"), v.get("unavailable"))
        print("   cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la $O | grep ${tag}_
