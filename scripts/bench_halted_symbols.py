#!/usr/bin/env python
"""A realistic panel: 20,000 symbols x 5,040 bars, device-resident, 1 % of the symbols (spread evenly) with a 3-bar trading halt
in close.  Kernel time of the fused suite with per-block dispatch only (PQB_COMPACT_NULLS=0: every block that holds a halted
symbol runs the null-aware kernel) and with symbol compaction (default: the halted symbols run in blocks of their own beside
the plain kernel on every original block)."""
import json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
S, NB = int(os.environ.get("PQB_BENCH_SYMBOLS", 20000)), 5040
p = pq.Panel(S, NB, engine=pq.get_engine(0))
p.fill_synthetic(seed=5, to_host=True)
prm = N.default_params()
tot, fused, nl = p.time_device(prm, warmup=2, iters=5)
base = tot / 5
print(json.dumps({"config": "no nulls", "symbols": S, "bars": NB, "total_ms": tot / 5, "launches": nl}), flush=True)
for frac in (0.001, 0.01, 0.05, 0.2):
    n = max(1, int(S * frac))
    ok = np.ones(NB, dtype=bool); ok[2000:2003] = False
    bits = np.packbits(ok, bitorder="little")
    for s in np.linspace(0, S - 1, n).astype(int):
        p.set_column(int(s), "close", np.ascontiguousarray(p.host_field("close")[int(s)]), validity=bits)
    p.upload()
    tot, fused, nl = p.time_device(prm, warmup=2, iters=5)
    print(json.dumps({"config": "%d symbols (%.1f %%) with a 3-bar halt in close" % (n, 100 * frac), "compaction": os.environ.get("PQB_COMPACT_NULLS", "1"),
                      "symbols": S, "bars": NB, "total_ms": tot / 5, "vs_no_nulls": tot / 5 / base, "launches": nl}), flush=True)
