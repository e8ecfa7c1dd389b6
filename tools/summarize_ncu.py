#!/usr/bin/env python
"""Turn gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and an optional .ncu-rep into
the tracked summaries under profiles/.  Usage: tools/summarize_ncu.py <tag> [launches.csv] [rep]"""
import collections
import csv
import io
import subprocess
import sys

tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv"
rep = sys.argv[3] if len(sys.argv) > 3 else None

lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
order = []
for row in csv.DictReader(io.StringIO("".join(lines))):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
    order.append((name, v))
tot = sum(v[1] for v in agg.values())
out = [f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)",
       "", f"command: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` -- {len(order)} launches, "
       f"{tot / 1e3:.2f} ms of kernel time", "",
       "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {v[0]} | {v[1] / 1e3:.2f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |")
if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "sm__cycles_elapsed.avg.per_second"]
    idx = [hdr.index(w) for w in want if w in hdr]
    out += ["", f"## `ncu --set full` capture ({rep.split('/')[-1]})", "",
            "| " + " | ".join(f"{hdr[i]} [{units[i]}]" for i in idx) + " |", "|" + "---|" * len(idx)]
    for r in rows[2:]:
        out.append("| " + " | ".join(r[i][:48].replace("void unnamed>::", "") for i in idx) + " |")
open(f"profiles/{tag}.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:30]))
