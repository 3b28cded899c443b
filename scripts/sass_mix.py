#!/usr/bin/env python
"""Static SASS op mix of the first (default-period, unmasked) steady tile body of suite_fused_kernel<32>.
usage: python scripts/sass_mix.py [lib.so]"""
import re, subprocess, sys, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "polars_quant_b200/libpqb200.so"
txt = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3pqb18suite_fused_kernelILi32EEEvNS_9SuiteArgsE", lib],
                     capture_output=True, text=True).stdout
L = []
for l in txt.splitlines():
    m = re.match(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?)\s*;", l)
    if m: L.append(m.group(2))
print("function instructions:", len(L))
stg = [i for i, l in enumerate(L) if "STG.E" in l and "256" in l and not l.startswith("@")]
# first cluster of 21 unpredicated 256-bit stores = steady<DEFP, !MASKED>
first = stg[:21]
lo = max(0, first[0] - 400); hi = first[-1] + 60
# tighten lo: last BAR.SYNC / SYNCS before first store
for i in range(first[0], lo, -1):
    if "SYNCS" in L[i] or "BAR.SYNC" in L[i]: lo = i; break
seg = L[lo:hi]
ops = collections.Counter()
for l in seg:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    ops[(m.group(2) if m else l).split(".")[0]] += 1
print("steady body ~", len(seg), "instructions (", lo, "..", hi, ")")
print(ops.most_common(30))
