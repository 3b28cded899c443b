#!/usr/bin/env python
"""Per-CUDA-line executed-instruction / stall-sample breakdown from an .ncu-rep (needs -lineinfo).
usage: python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
agg = []
for r in rows[h + 1:]:
    if len(r) > iE and r[0].strip().isdigit():
        try: agg.append((int(r[iE]), int(r[iS]), int(r[0]), r[1].strip()[:100]))
        except ValueError: pass
tot = sum(a[0] for a in agg) or 1; ts = sum(a[1] for a in agg) or 1
print("total warp-instructions", tot, "samples", ts)
for n, s, ln, t in sorted(agg, reverse=True)[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*s/ts:5.1f}% samp  L{ln}: {t}")
