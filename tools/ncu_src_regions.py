"""Aggregate an `ncu --page source --csv` dump into regions of SASS delimited by BAR.SYNC, with per-region stall mix.
usage: python tools/ncu_src_regions.py src.csv [min_samples]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
minS = int(sys.argv[2]) if len(sys.argv) > 2 else 200
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
reg = []
cur = dict(start=0, n=0, samples=0, st=collections.Counter(), ops=collections.Counter(), first="", execd=0)
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    s = int(r[ix["# Samples"]] or 0)
    cur["n"] += 1; cur["samples"] += s
    cur["execd"] += int(r[ix["Instructions Executed"]] or 0)
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    cur["ops"][op.split(".")[0]] += 1
    for n in stalls:
        v = int(r[ix[n]] or 0)
        if v:
            cur["st"][n] += v
    if "BAR.SYNC" in src or "EXIT" in src:
        cur["end"] = k
        reg.append(cur)
        cur = dict(start=k + 1, n=0, samples=0, st=collections.Counter(), ops=collections.Counter(), first="", execd=0)
tot = sum(r["samples"] for r in reg)
print("total samples", tot)
for r in reg:
    if r["samples"] < minS:
        continue
    top = ", ".join(f"{k[6:]}={v}" for k, v in r["st"].most_common(5))
    ops = ", ".join(f"{k}:{v}" for k, v in r["ops"].most_common(6))
    print(f"lines {r['start']:5d}-{r['end']:5d} ({r['n']:4d} instr, exec {r['execd']:9d}) samples {r['samples']:7d} ({100*r['samples']/tot:5.1f}%)  {top}  | {ops}")
