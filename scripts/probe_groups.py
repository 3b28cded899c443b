#!/usr/bin/env python
"""Kernel-only time of each optional group alone and of all of them on S x 5,040 (argv[1] = S, default 50,000)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
S = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
p = pq.Panel(S, 5_040, engine=pq.get_engine(0), outputs_mask=(1 << N.N_OUTPUTS) - 1, host_staging=False)
p.fill_synthetic(seed=7)
for name, bit in list(N.IND_EXTRA.items()) + [("every optional group", sum(N.IND_EXTRA.values()))]:
    tot, fused, nl = p.time_device(N.default_params(indicators=bit), warmup=1, iters=3)
    print(f"{S:6d} x 5040 {name:22s} {fused / 3:8.3f} ms", flush=True)
