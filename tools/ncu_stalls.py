"""Summarise an ncu report here (no GPU): headline metrics + warp-stall samples aggregated by opcode.
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [kernel-index]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
d = dict(zip(hdr, vals))
for k in ("Kernel Name", "gpu__time_duration.sum", "sm__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"):
    if k in d:
        print(f"{k:70s} {d[k][:100]}")
for k, v in d.items():
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        print(f"  {k.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {float(v):.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
tot = collections.Counter(); byop = collections.defaultdict(collections.Counter); ex = collections.Counter()
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None: break
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or len(r) != len(hdr): continue
    s = r[ix["Source"]].split()
    op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    ex[op] += int(r[ix["Instructions Executed"]] or 0)
    for st in stalls:
        v = int(r[ix[st]] or 0); tot[st] += v; byop[op][st] += v
n = sum(tot.values())
print("samples", n, {k: f"{v/n:.1%}" for k, v in tot.most_common(8)})
for op, c in sorted(byop.items(), key=lambda kv: -sum(kv[1].values()))[:10]:
    print(f"  {op:8s} samples {sum(c.values()):6d} executed {ex[op]:10d}  {dict(c.most_common(5))}")
