#!/usr/bin/env python
"""Samples per role region (code between consecutive per-stage mbarrier waits) from an ncu source CSV."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
S = lambda r, h: int(r[ix[h]])
waits = [i for i, r in enumerate(data) if 'SYNCS.PHASECHK.TRANS64.TRYWAIT' in r[1]]
print("try_wait sites at", waits)
bounds = waits + [len(data)]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for k in range(len(waits)):
    a, b = bounds[k], bounds[k + 1]
    rs = data[a:b]
    n = sum(S(r, '# Samples') for r in rs)
    st = {h[6:]: sum(S(r, h) for r in rs) for h in stalls}
    st = {k2: v for k2, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > 0.04 * n}
    ops = collections.Counter()
    for r in rs:
        t = r[1].split()
        if not t: continue
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] += S(r, 'Instructions Executed')
    hot = max(S(r, 'Instructions Executed') for r in rs)
    print(f"region {a}-{b} instrs {b-a} samples {n} exec_total {sum(ops.values())} MUFU {ops['MUFU']} DFMA {ops['DFMA']} CALL {ops['CALL']}")
    print("     ", st)
