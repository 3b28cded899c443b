#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list (csv written by
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu`): DRAM bytes per launch of the dominant kernel of each workload,
stamped with the commit the capture was made at (the GPU box has no .git: the caller passes it).
usage: python scripts/ncu_traffic.py launches.csv <commit> [summary.txt]"""
import collections
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
src, commit = sys.argv[1], sys.argv[2]
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
per = collections.OrderedDict()          # (launch id) -> {kernel, grid, metrics}
for r in rd:
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    unit = r.get("Metric Unit", "")
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "second": 1e3, "s": 1e3}.get(unit, 1.0)
    e = per.setdefault(r["ID"], {"kernel": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
    e[r["Metric Name"]] = v * mult

launches = list(per.values())
def short(k):
    return k.split("(")[0]

# summary per kernel name: launches, total ms, share, DRAM bytes per launch (mean / max)
agg = collections.OrderedDict()
for e in launches:
    a = agg.setdefault(short(e["kernel"]), {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0, "max_ms": 0.0, "grid": e["grid"]})
    a["n"] += 1
    a["ms"] += e.get("gpu__time_duration.sum", 0.0)
    a["max_ms"] = max(a["max_ms"], e.get("gpu__time_duration.sum", 0.0))
    a["rd"] += e.get("dram__bytes_read.sum", 0.0)
    a["wr"] += e.get("dram__bytes_write.sum", 0.0)
tot = sum(a["ms"] for a in agg.values()) or 1.0
out = ["# launch list summary of %s, captured at commit %s" % (Path(src).name, commit),
       "# ncu times are cold-cache and serialised: the SHARE of a kernel counts, not the absolute",
       "%-70s %6s %10s %7s %10s %14s %14s" % ("kernel", "n", "total ms", "share", "max ms", "DRAM rd/launch", "DRAM wr/launch")]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    out.append("%-70s %6d %10.3f %6.1f%% %10.3f %14.4g %14.4g" % (k[:70], a["n"], a["ms"], 100 * a["ms"] / tot, a["max_ms"], a["rd"] / a["n"], a["wr"] / a["n"]))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 3:
    Path(sys.argv[3]).write_text(txt + "\n")

# the headline kernel: the longest suite_fused launch = one launch over the 50,000 x 5,040 panel
recs = []
big = [e for e in launches if "suite_fused_kernel" in e["kernel"]]
if big:
    e = max(big, key=lambda e: e.get("gpu__time_duration.sum", 0.0))
    recs.append({"symbols": 50000, "bars": 5040, "kernel": short(e["kernel"]), "grid": e["grid"],
                 "dram_bytes_read": e.get("dram__bytes_read.sum"), "dram_bytes_write": e.get("dram__bytes_write.sum"),
                 "dram_bytes_per_launch": e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0),
                 "algorithmic_bytes_per_launch": 200 * 50000 * 5040, "ncu_ms": e.get("gpu__time_duration.sum"),
                 "captured_at_commit": commit,
                 "source": "profiles/%s (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over bench.py --steps 2 --warmup 1 --no-e2e --no-cpu)" % Path(src).name})
ks = [e for e in launches if "window_suite_kernel" in e["kernel"]]
if ks:
    e = max(ks, key=lambda e: e.get("gpu__time_duration.sum", 0.0))
    recs.append({"workload": "c5", "symbols_": 10000, "bars_": 5040, "kernel": short(e["kernel"]), "grid": e["grid"],
                 "dram_bytes_read": e.get("dram__bytes_read.sum"), "dram_bytes_write": e.get("dram__bytes_write.sum"),
                 "dram_bytes_per_launch": e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0),
                 "algorithmic_bytes_per_pass": 248 * 10000 * 5040, "ncu_ms": e.get("gpu__time_duration.sum"), "captured_at_commit": commit})
# config 3: one pass = seed + local + carry + final; per-launch means of the four kernels, summed
lr = {k: a for k, a in agg.items() if "lr_" in k}
if lr:
    recs.append({"workload": "c3", "symbols_": 500, "bars_": 1000000, "kernel": " + ".join(sorted(lr)),
                 "dram_bytes_read": sum(a["rd"] / a["n"] for a in lr.values()), "dram_bytes_write": sum(a["wr"] / a["n"] for a in lr.values()),
                 "dram_bytes_per_launch": sum((a["rd"] + a["wr"]) / a["n"] for a in lr.values()),
                 "algorithmic_bytes_per_pass": 64 * 500 * 1000000, "ncu_ms": sum(a["ms"] / a["n"] for a in lr.values()),
                 "captured_at_commit": commit})
if recs and recs[0].get("symbols") == 50000:
    (ROOT / "gpurun_out" / "traffic.json").write_text(json.dumps(recs, indent=1) + "\n")
    print("wrote gpurun_out/traffic.json (copy to profiles/traffic.json)")
