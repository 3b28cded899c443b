#!/usr/bin/env python
"""Per-output bit-exactness report of the CUDA suite against the C oracle (GPU box only)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import polars_quant_b200 as pq
import synth
from oracle import pqo

S, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (300, 2520)
d = synth.ohlcv(S, N, seed=11)
p = pq.Panel(S, N)
p.set_fields(d["close"], d["high"], d["low"], d["volume"])
res = p.compute()
out, ok, _ = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"])
for j, name in enumerate(pqo.OUTPUT_NAMES):
    gv, gok = res[name]
    vbad = int((gok != ok[j]).sum())
    m = ok[j] & gok
    a, b = gv[m], out[j][m]
    same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
    nb = int((~same).sum())
    worst = float(np.max(np.abs(a[~same] - b[~same]) / np.maximum(np.abs(b[~same]), 1e-300))) if nb else 0.0
    first = ""
    if nb:
        idx = np.argwhere(m & ~((gv.view(np.uint64) == out[j].view(np.uint64)) | (np.isnan(gv) & np.isnan(out[j]))))[0]
        first = f" first at {tuple(int(x) for x in idx)} gpu={gv[tuple(idx)]!r} ref={out[j][tuple(idx)]!r}"
    print(f"{name:12s} validity_mismatch={vbad:6d} value_bits_differ={nb:8d} of {int(m.sum()):9d} worst_rel={worst:.2e}{first}")
