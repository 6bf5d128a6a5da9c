"""CPU: the JSON line of `bench.py --impl reference` (the one arm that runs without a GPU) carries the keys the driver
reads, and the B200 arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def run(*args):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_json_contract():
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "GP/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("homogenized GPs/sec") and d["dtype"] == "f64" and d["scaling"] == "weak"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"] and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "GP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return   # on a GPU box the arm runs for real (covered by the driver)
    r = run("--steps", "1", "--warmup", "0", "--ngp", "2", "--no-cpu-baseline")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
