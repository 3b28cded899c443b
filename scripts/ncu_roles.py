#!/usr/bin/env python
"""Per steady-loop (= per role) stall-sample breakdown from an ncu source-page CSV export
(ncu -i rep --page source --csv --print-source sass > src.csv).  usage: ncu_roles.py src.csv [min_exec]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
S = lambda r, h: int(r[ix[h]])
cnt = collections.Counter(S(r, 'Instructions Executed') for r in data)
hot = max((c for c in cnt if c > 0), key=lambda c: c * cnt[c])
lo = hot * 0.8
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
segs = []; cur = None
for i, r in enumerate(data):
    if S(r, 'Instructions Executed') >= lo:
        if cur is None or i - cur[1] > 12: cur = [i, i]; segs.append(cur)
        cur[1] = i
tot = sum(S(r, '# Samples') for r in data)
print("hot exec count", hot, "total samples", tot)
for a, b in segs:
    rs = data[a:b + 1]
    n = sum(S(r, '# Samples') for r in rs)
    ops = collections.Counter(r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1] for r in rs)
    st = {h[6:]: sum(S(r, h) for r in rs) for h in stalls}
    st = {k: v for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > 0.03 * n}
    sig = [o for o in ('MUFU.RCP64H', 'MUFU.RSQ64H') if ops[o]]
    print(f"seg {a}-{b} n_instr {b-a+1} samples {n} ({100*n/tot:.1f}%) DFMA {ops['DFMA']} DADD {ops['DADD']} DMUL {ops['DMUL']} LDS {ops['LDS.64']} STS {ops['STS.64']} STG {sum(v for k,v in ops.items() if k.startswith('STG'))} RCP {ops['MUFU.RCP64H']} RSQ {ops['MUFU.RSQ64H']} BRA {sum(v for k,v in ops.items() if k.startswith('BRA'))}")
    print("     ", st)
