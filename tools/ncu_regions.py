"""Per-region view of an ncu source page (no GPU): the SASS of a kernel is cut at its barriers (BAR.SYNC / UCGABAR_WAIT)
and, per region, the stall samples, executed instructions, shared-memory wavefronts (ideal vs actual) and the dominant
opcodes are printed.   python tools/ncu_regions.py report.ncu-rep"""
import collections, csv, io, subprocess, sys
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
regions = []
cur = None
def new():
    return dict(samples=0, inst=0, wf=0, wfi=0, ops=collections.Counter(), stalls=collections.Counter(), first=None, dfma=0, lds=0, local=0)
cur = new()
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None: break
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) != len(hdr): continue
    s = r[ix["Source"]].split()
    op = (s[1] if s[0].startswith("@") else s[0])
    base = op.split(".")[0]
    n = int(r[ix["# Samples"]] or 0)
    cur["samples"] += n
    cur["inst"] += int(r[ix["Instructions Executed"]] or 0)
    cur["wf"] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    cur["wfi"] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    cur["ops"][base] += n
    if base == "DFMA": cur["dfma"] += int(r[ix["Instructions Executed"]] or 0)
    if base in ("LDL", "STL"): cur["local"] += int(r[ix["Instructions Executed"]] or 0)
    for st in stalls:
        cur["stalls"][st] += int(r[ix[st]] or 0)
    if base in ("BAR", "UCGABAR_WAIT", "EXIT"):
        cur["end"] = op
        regions.append(cur); cur = new()
regions.append(cur)
tot = sum(x["samples"] for x in regions)
print("total samples", tot)
for i, x in enumerate(regions):
    if x["samples"] < tot * 0.002: continue
    print(f"region {i:2d} ends {x.get('end','-'):14s} samples {x['samples']:7d} ({x['samples']/tot:5.1%}) inst {x['inst']:11d} DFMA {x['dfma']:10d} "
          f"local {x['local']:8d} smem wf {x['wf']:10d} ideal {x['wfi']:10d}  top ops {dict(x['ops'].most_common(4))}  stalls "
          f"{ {k[6:]: v for k, v in x['stalls'].most_common(4)} }")
