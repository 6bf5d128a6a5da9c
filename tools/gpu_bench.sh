#!/bin/bash
# gpurun --timeout 1500 -- 'bash tools/gpu_bench.sh <tag> [bench args]' : the default bench line (both arms) under gpurun_out/
tag=${1:-b}; shift
mkdir -p gpurun_out
timeout 1200 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -c 1500 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_bench.json") if l.startswith("{")][-1])
    print("value", d["value"], d["unit"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "roof", d["roofline"].get("fp64", d["roofline"]).get("frac"), d["roofline"]["share_of_step"], "cpu", d.get("cpu_baseline", {}).get("value"))
    for k, w in d.get("workloads", {}).items():
        print(" ", k, w.get("value"), w.get("unit"), "e2e", w.get("e2e", {}).get("value"), "wall", w.get("bench_wall_s"), "cpu", w.get("cpu_baseline"), w.get("unavailable"), w.get("vs_reference_fixture"), w.get("cg_its"))
except Exception as ex:
    print("parse failed", ex)
PY
