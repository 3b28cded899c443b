#!/usr/bin/env python
"""Kernel time of the null-aware suite kernel against the plain one on the same 8,192 x 5,040 panel (GPU box): the panel
switches to the null-aware kernel as soon as ONE symbol has an interior null (a trading halt)."""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
S, NB = 8192, 5040
p = pq.Panel(S, NB, engine=pq.get_engine(0))
p.fill_synthetic(seed=5, to_host=True)
prm = N.default_params()
tot, fused, nl = p.time_device(prm, warmup=2, iters=5)
print(json.dumps({"config": "plain kernel, no nulls", "symbols": S, "bars": NB, "kernel_ms": fused / 5, "total_ms": tot / 5, "launches": nl}))
for n_halted in (1, 82, 8192):
    ok = np.ones(NB, dtype=bool); ok[2000:2003] = False
    bits = np.packbits(ok, bitorder="little")
    step = max(1, S // n_halted)
    for s in range(0, S, step):
        p.set_column(s, "close", np.ascontiguousarray(p.host_field("close")[s]), validity=bits)
    p.upload()
    tot, fused, nl = p.time_device(prm, warmup=2, iters=5)
    print(json.dumps({"config": "null-aware kernel, %d symbols with a 3-bar halt in close" % len(range(0, S, step)), "symbols": S, "bars": NB,
                      "kernel_ms": fused / 5, "total_ms": tot / 5, "launches": nl}))
