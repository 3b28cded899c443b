#!/usr/bin/env python
"""Innermost loops (backward branches) of one kernel in a .so: size and op mix.  usage: sass_loops.py lib.so [kernel-substring]"""
import re, subprocess, sys, collections
lib = sys.argv[1]; key = sys.argv[2] if len(sys.argv) > 2 else "suite_fused_kernelILb1ELb0"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins = []; on = False
for l in txt.splitlines():
    if "Function :" in l: on = key in l
    if not on: continue
    m = re.match(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?)\s*;", l)
    if m: ins.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(ins)}
print("instructions:", len(ins))
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.\S+)?\s+(?:\S+,\s+)?0x([0-9a-f]+)$", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr: loops.append((addr[tgt], i))
# keep innermost-ish loops with stores
for lo, hi in loops:
    body = [t for _, t in ins[lo:hi + 1]]
    op = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in body)
    stg = sum(v for k, v in op.items() if k.startswith("STG"))
    if hi - lo < 12: continue
    print(f"loop {ins[lo][0]:#x}-{ins[hi][0]:#x} n={hi-lo+1:4d} DFMA {op['DFMA']:3d} DADD {op['DADD']:3d} DMUL {op['DMUL']:3d} MUFU {op['MUFU.RCP64H']+op['MUFU.RSQ64H']} LDS {op['LDS.64']:2d} STS {op['STS.64']:2d} STG {stg:2d} BRA {sum(v for k,v in op.items() if k.startswith('BRA')):2d} CALL {sum(v for k,v in op.items() if k.startswith('CALL'))} LDL/STL {op['LDL']+op['STL']+op['LDL.64']+op['STL.64']}")
