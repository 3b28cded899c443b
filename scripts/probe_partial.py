#!/usr/bin/env python
"""Kernel-only timing of partial suites (the BASE kernel) at two panel sizes."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import polars_quant_b200 as pq
from polars_quant_b200 import _native as NV
eng = pq.get_engine(0)
G = {n: 1 << i for i, n in enumerate(("sma", "ema", "tema", "trima", "bbands", "macd", "rsi", "trange", "atr", "natr", "obv", "ad", "kdj", "willr", "midprice"))}
sets = {"ema": G["ema"], "rsi": G["rsi"], "bbands": G["bbands"], "kdj+atr": G["kdj"] | G["atr"], "sma+ema+rsi+macd+bbands": G["sma"] | G["ema"] | G["rsi"] | G["macd"] | G["bbands"],
        "all but kdj": NV.IND_ALL & ~G["kdj"]}
for S, N in ((4736, 2520), (50_000, 5_040)):
    p = pq.Panel(S, N, engine=eng, host_staging=False)
    p.fill_synthetic(seed=1, sigma=0.02)
    for name, m in sets.items():
        tot, fused, nl = p.time_device(NV.default_params(indicators=m), warmup=2, iters=5)
        print(f"{S:6d} x {N:5d} {name:28s} {fused / 5:8.3f} ms")
    p.close()
