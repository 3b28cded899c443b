#!/usr/bin/env python
"""Kernel-only timing of the fused suite over a list of panel shapes: S1xN1 S2xN2 ... (GPU box)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import polars_quant_b200 as pq
from polars_quant_b200 import _native as NV
eng = pq.get_engine(0)
for spec in sys.argv[1:]:
    S, N = (int(x) for x in spec.split("x"))
    p = pq.Panel(S, N, engine=eng, host_staging=False)
    p.fill_synthetic(seed=1, sigma=0.02)
    tot, fused, nl = p.time_device(NV.default_params(), warmup=3, iters=10)
    ms = fused / 10
    print(f"{S:6d} x {N:6d}: {ms:8.3f} ms  {200.0*S*N/ms/1e6:7.0f} GB/s  {S*N/ms/1e6:8.2f} G symbol-bars/s  cyc/bar/block {ms*1e-3*1.965e9/N/max(1,-(-S//32)//148):.0f}")
    p.close()
