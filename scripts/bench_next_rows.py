#!/usr/bin/env python
"""Measurements for the SURVEY.md 8(f) rows that are not on the headline path (GPU box):
  optional groups : 10,000 x 5,040, the DM family alone (3 in + 6 out = 72 B per symbol-bar) and the whole 41-output suite
                    (4 in + 41 out = 360 B), kernel-only, device-resident
  signals         : 20,000 x 5,040, the crossover kernel alone (5 planes read + 3 int8 written = 43 B per symbol-bar)
  info            : 50,000 x 5,040, Panel.info() wall clock (kernel + 5 MB device->host)
  c5 split        : 10,000 x 5,040 WILLR(p) + MIDPRICE(p) through a time-split panel (pure windows: bit-exact)
  wide            : a 2,000-symbol x 2,520-bar Arrow table through WidePanel.suite() / .candles() / .info(), wall clock of
                    the whole call (Arrow columns -> pinned staging -> GPU -> zero-copy Arrow results)"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import polars_quant_b200 as pq
from polars_quant_b200 import _native as N

NAMES = N.OUTPUT_NAMES
peak = 6550.0
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
eng = pq.get_engine(0)
out = []
which = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "groups,signals,info,c5split,wide"


def rec(**kw):
    print(json.dumps(kw), flush=True)
    out.append(kw)


def kernel_rec(tag, S, NB, ms, bps, **extra):
    rec(config=tag, symbols=S, bars=NB, kernel_ms=ms, algorithmic_bytes_per_symbol_bar=bps, achieved_gbs=bps * S * NB / ms / 1e6,
        frac=bps * S * NB / ms / 1e6 / peak, symbol_bars_per_s=S * NB / ms * 1e3, **extra)


if "groups" in which:
    # bytes per symbol-bar: 8 x (input planes the groups read + outputs they write); the suite + optional groups run as
    # two launches that both read their inputs
    OUTS = {"midpoint": 1, "adosc": 1, "mom": 1, "roc": 4, "cmo": 1, "mfi": 1, "cci": 1, "dm": 6, "trix": 1, "ultosc": 1, "aroon": 2, "donchian": 2}
    INS = {"midpoint": 1, "adosc": 4, "mom": 1, "roc": 1, "cmo": 1, "mfi": 4, "cci": 3, "dm": 3, "trix": 1, "ultosc": 3, "aroon": 2, "donchian": 2}
    for S in (10_000, 50_000):
        NB = 5_040
        p = pq.Panel(S, NB, engine=eng, outputs_mask=(1 << N.N_OUTPUTS) - 1, host_staging=False)
        p.fill_synthetic(seed=7)
        if S == 10_000:
            for name, bit in N.IND_EXTRA.items():
                tot, fused, nl = p.time_device(N.default_params(indicators=bit), warmup=1, iters=3)
                kernel_rec("optional group alone: " + name, S, NB, fused / 3, 8 * (INS[name] + OUTS[name]))
        tot, fused, nl = p.time_device(N.default_params(indicators=N.IND_ALL), warmup=1, iters=3)
        kernel_rec("the 15-indicator suite (same panel)", S, NB, fused / 3, 200)
        tot, fused, nl = p.time_device(N.default_params(indicators=N.IND_ALL | N.IND_EXTRA["mom"]), warmup=1, iters=3)
        kernel_rec("suite + MOM (two launches)", S, NB, fused / 3, 200 + 16)
        allg = N.IND_ALL | sum(N.IND_EXTRA.values())
        tot, fused, nl = p.time_device(N.default_params(indicators=allg), warmup=1, iters=3)
        kernel_rec("every group: %d outputs (two launches)" % N.N_OUTPUTS, S, NB, fused / 3, 8 * (4 + 21) + 8 * (4 + N.N_OUTPUTS - 21))
        p.close()

if "signals" in which:
    S, NB = 20_000, 5_040
    om = sum(1 << NAMES.index(o) for o in ("macd", "macd_signal", "kdj_k", "kdj_d", "rsi"))
    p = pq.Panel(S, NB, engine=eng, outputs_mask=om, host_staging=False)
    p.fill_synthetic(seed=11)
    p.run(N.default_params(indicators=N.IND["macd"] | N.IND["kdj"] | N.IND["rsi"]))
    ms = C.c_float()
    N.check(N.lib().pqb_signals_time(p._h, 30.0, 70.0, 3, 20, C.byref(ms)))
    kernel_rec("crossover signals (macd, kdj, rsi)", S, NB, ms.value / 20, 43)
    p.close()

if "info" in which:
    S, NB = 50_000, 5_040
    p = pq.Panel(S, NB, engine=eng, outputs_mask=1, host_staging=False)
    p.fill_synthetic(seed=13)
    p.info()
    t0 = time.perf_counter()
    for _ in range(10):
        p.info()
    ms = (time.perf_counter() - t0) * 100.0
    rec(config="Panel.info(): 13 last-row reductions per symbol", symbols=S, bars=NB, wall_ms_per_call=ms,
        bytes_read=S * 21 * 4 * 8, note="kernel + 13 x n_symbols doubles device->host; reads the last 21 bars only")
    p.close()

if "c5split" in which:
    S, NB = 10_000, 5_040
    om = sum(1 << NAMES.index(o) for o in ("willr", "midprice"))
    for w, chunks in ((20, 4), (20, 8), (55, 8), (250, 4)):
        prm = N.default_params(indicators=N.IND["willr"] | N.IND["midprice"], willr_period=w, midprice_period=w)
        W = pq.SplitPanel.required_warmup(prm)
        sp = pq.SplitPanel(S, NB, chunks=chunks, warmup=W, engine=eng, fields_mask=0b0111, outputs_mask=om, host_staging=False)
        sp.fill_synthetic(seed=55, sigma=0.02)
        tot, fused, nl = sp.time_device(prm, warmup=2, iters=5)
        kernel_rec(f"c5 split willr({w})+midprice({w})", S, NB, fused / 5, 40, chunks=chunks, warmup=sp.warmup,
                   virtual_symbols=sp.virtual_symbols, virtual_bars=sp.virtual_bars)
        sp.close()

if "wide" in which:
    import pyarrow as pa
    import synth
    from polars_quant_b200 import wide
    S, NB = 2_000, 2_520
    d = synth.ohlcv(S, NB, seed=17)
    cols, names = [pa.array(np.arange(NB, dtype=np.int32))], ["date"]
    for s in range(S):
        for f in ("open", "high", "low", "close", "volume"):
            cols.append(pa.array(d[f][s]))
            names.append("S%05d_%s" % (s, f))
    table = pa.table(cols, names=names)
    wp = wide.WidePanel(table, engine=eng)
    for name, fn in (("suite", wp.suite), ("candles", wp.candles), ("info", wp.info)):
        fn()
        t0 = time.perf_counter()
        res = fn()
        dt = time.perf_counter() - t0
        rec(config="WidePanel.%s(): Arrow table in, Arrow table out" % name, symbols=S, bars=NB, wall_ms=dt * 1e3,
            result_columns=res.num_columns, symbol_bars_per_s=S * NB / dt)
        del res

if "--json" in sys.argv:
    Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps(out, indent=1))
