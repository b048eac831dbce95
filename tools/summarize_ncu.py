"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/summarize_ncu.py <round-tag> <launches.csv> [<prof.ncu-rep>]"""
import collections, csv, json, re, subprocess, sys
from pathlib import Path

tag, launches = sys.argv[1], sys.argv[2]
rep = sys.argv[3] if len(sys.argv) > 3 else None
out = Path("profiles"); out.mkdir(exist_ok=True)

rows = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, vi, ni = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[ni] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("bvb::", "")
    agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
lines = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` on `python tools/run_sweeps.py 12`",
         "# (cold-cache, serialised: compare SHARES, not absolutes)", "",
         f"| kernel | launches | avg us | share of step |", "|---|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"| {k} | {len(v)} | {sum(v)/len(v)*1e-3:.1f} | {100*sum(v)/tot:.1f}% |")
(out / f"launches_{tag}.md").write_text("\n".join(lines) + "\n")

if rep is None:
    sys.exit(0)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
dmma = [i for i, x in enumerate(h) if "subpipe_dmma_cycles_active.avg" in x]
summary = {}
for r in rr[2:]:
    name = re.sub(r"\(.*", "", r[h.index("Kernel Name")]).replace("void ", "").replace("bvb::", "")
    name = re.sub(r"<.*", "", name)
    d = {}
    for w in want:
        if w in h:
            try:
                d[w] = float(r[h.index(w)].replace(",", ""))
                d[w + ".unit"] = units[h.index(w)]
            except ValueError:
                pass
    if dmma:
        d["smsp__pipe_tensor_subpipe_dmma_cycles_active.avg"] = float(r[dmma[0]].replace(",", ""))
    def tobytes(key):
        v, u = d.get(key), d.get(key + ".unit", "byte")
        return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rd, wr = tobytes("dram__bytes_read.sum"), tobytes("dram__bytes_write.sum")
    if rd is not None and wr is not None:
        d["dram_bytes_per_launch"] = rd + wr
    if "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg" in d and "sm__cycles_elapsed.max" in d:
        d["dmma_pipe_active_frac_of_elapsed"] = d["smsp__pipe_tensor_subpipe_dmma_cycles_active.avg"] / d["sm__cycles_elapsed.max"]
    summary.setdefault(name, d)
(out / f"ncu_summary_{tag}.json").write_text(json.dumps(summary, indent=1))
(out / "ncu_summary.json").write_text(json.dumps(summary, indent=1))
for k, d in summary.items():
    print(k, {x: d[x] for x in ("gpu__time_duration.sum", "dram_bytes_per_launch", "dmma_pipe_active_frac_of_elapsed", "launch__registers_per_thread") if x in d})
