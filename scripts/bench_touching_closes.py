#!/usr/bin/env python
"""Does real-looking data cost the fused suite more than the synthetic panel?  On real bars close == high (or == low) on a
few percent of the rows, which makes WILLR's / STOCH's numerators exactly zero now and then: a zero numerator fails the
fast path of an IEEE division (the AROON finding).  8,192 x 5,040, device-resident, CUDA events."""
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
tot, fused, nl = p.time_device(prm, warmup=2, iters=10)
print(json.dumps({"config": "synthetic panel", "kernel_ms": fused / 10}))
rng = np.random.default_rng(1)
for frac in (0.02, 0.10):
    c, h, l = p.host_field("close"), p.host_field("high"), p.host_field("low")
    m = rng.random((S, NB)) < frac
    h[:, :NB][m] = c[:, :NB][m]                       # close at the high of the bar
    m = rng.random((S, NB)) < frac
    l[:, :NB][m] = c[:, :NB][m]                       # close at the low of the bar
    p.upload()
    tot, fused, nl = p.time_device(prm, warmup=2, iters=10)
    print(json.dumps({"config": "close == high on %.0f%% and close == low on %.0f%% of the bars" % (frac * 100, frac * 100), "kernel_ms": fused / 10}))
# the same without any bar where high == low == close: only the highs touch
p.fill_synthetic(seed=5, to_host=True)
c, h = p.host_field("close"), p.host_field("high")
m = rng.random((S, NB)) < 0.10
h[:, :NB][m] = c[:, :NB][m]
p.upload()
tot, fused, nl = p.time_device(prm, warmup=2, iters=10)
print(json.dumps({"config": "close == high on 10% of the bars, lows untouched", "kernel_ms": fused / 10}))
