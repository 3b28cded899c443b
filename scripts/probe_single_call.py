#!/usr/bin/env python
"""Where a single-column call's time goes: the C-ABI entry point (pqb_ema / pqb_macd / pqb_bbands, what a Rust host calls) next to
the same call through the Python harness of the polars plugin symbol (ctypes + pyarrow export / import around it)."""
import ctypes as C, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, pyarrow as pa
import synth
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N, talib
L = N.lib(); eng = pq.get_engine(0)
res = {}
for n in (252, 2520, 25200):
    d = synth.ohlcv(1, n, seed=1)
    a = np.ascontiguousarray(d["close"][0]); ac = pa.array(a)
    c = N.Col(a.ctypes.data, None, 0, n)
    vals = [np.empty(n) for _ in range(3)]; bits = [np.zeros((n + 7) // 8, np.uint8) for _ in range(3)]
    oc = [N.OutCol(v.ctypes.data, b.ctypes.data) for v, b in zip(vals, bits)]
    calls = {"ema": (lambda: L.pqb_ema(eng._h, C.byref(c), 30, C.byref(oc[0])), lambda: talib.EMA(ac)),
             "rsi": (lambda: L.pqb_rsi(eng._h, C.byref(c), 14, C.byref(oc[0])), lambda: talib.RSI(ac)),
             "macd": (lambda: L.pqb_macd(eng._h, C.byref(c), 12, 26, 9, C.byref(oc[0]), C.byref(oc[1]), C.byref(oc[2])), lambda: talib.MACD(ac)),
             "bbands": (lambda: L.pqb_bbands(eng._h, C.byref(c), 20, C.c_double(2.0), C.c_double(2.0), C.byref(oc[0]), C.byref(oc[1]), C.byref(oc[2])), lambda: talib.BBANDS(ac))}
    for name, (raw, py) in calls.items():
        out = {}
        for tag, f in (("c_abi_us", raw), ("python_harness_us", py)):
            for _ in range(20): f()
            t0 = time.perf_counter(); k = 300
            for _ in range(k): f()
            out[tag] = round((time.perf_counter() - t0) / k * 1e6, 1)
        res["%s_%d" % (name, n)] = out
print(json.dumps(res))
