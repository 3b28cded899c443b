#!/usr/bin/env python
"""BASELINE config 1: SMA / EMA / RSI / MACD / BBANDS on 1 symbol x 252 daily bars through the Python API (the polars
plugin symbols), per-call latency on the GPU box, with the C oracle (the reference's loops) timed beside it."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, pyarrow as pa
import synth
from oracle import pqo
from polars_quant_b200 import talib
d = synth.ohlcv(1, 252, seed=1)
c = d["close"][0]; ac = pa.array(c)
calls = {"SMA": (lambda: talib.SMA(ac), lambda: pqo.sma(c, 30)), "EMA": (lambda: talib.EMA(ac), lambda: pqo.ema(c, 30)),
         "RSI": (lambda: talib.RSI(ac), lambda: pqo.rsi(c, 14)), "MACD": (lambda: talib.MACD(ac), lambda: pqo.macd(c)),
         "BBANDS": (lambda: talib.BBANDS(ac), lambda: pqo.bbands(c))}
out = {}
for name, (g, o) in calls.items():
    for f in (g, o): f()
    t0 = time.perf_counter(); n = 300
    for _ in range(n): g()
    tg = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n): o()
    to = (time.perf_counter() - t0) / n
    out[name] = {"gpu_plugin_call_us": tg * 1e6, "cpu_oracle_call_us_incl_ctypes": to * 1e6}
print(json.dumps({"config": "BASELINE config 1: 1 symbol x 252 bars via the Python API", "per_call": out}))
