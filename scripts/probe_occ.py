#!/usr/bin/env python
"""Kernel-only time of one partial suite at 50,000 x 5,040 (argv[1] = ema | rsi | bbands); occupancy knobs come from the environment."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import polars_quant_b200 as pq
from polars_quant_b200 import _native as NV
p = pq.Panel(50_000, 5_040, engine=pq.get_engine(0), host_staging=False)
p.fill_synthetic(seed=1, sigma=0.02)
for name in sys.argv[1:]:
    tot, fused, nl = p.time_device(NV.default_params(indicators=sum(NV.IND[g] for g in name.split('+'))), warmup=2, iters=5)
    print(f"{name:24s} {fused / 5:8.3f} ms", flush=True)
