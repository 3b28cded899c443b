"""Generates tests/golden/*.npz from the pure-Python restatement oracle/ref_py.py.

Run from the repo root:  python tests/golden/make_golden.py
The reference itself cannot run here, so these vectors pin the C
oracle and the CUDA path against an independent restatement of the Rust text, not against
reference output ("parity unpinned").  Files are small (1 x 252 / 1 x 300 columns).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import ref_py as R  # noqa: E402
import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def run_all(c, h, l, v, tag, arrays):
    """c/h/l/v: list[float|None]; stores every function's outputs under '<tag>/<fn>[/k]'."""
    def put(name, cols):
        if not isinstance(cols, tuple):
            cols = (cols,)
        for k, col in enumerate(cols):
            vals, ok = synth.from_opt(col)
            arrays[f"{tag}/{name}/{k}/v"] = vals
            arrays[f"{tag}/{name}/{k}/ok"] = ok

    def attempt(name, fn):
        try:
            put(name, fn())
        except R.RefError:
            arrays[f"{tag}/{name}/err"] = np.array([1])

    for p in (1, 2, 5, 30):
        attempt(f"sma_{p}", lambda: R.calc_sma(c, p))
        attempt(f"ema_{p}", lambda: R.calc_ema(c, p))
        attempt(f"tema_{p}", lambda: R.calc_tema(c, p))
        attempt(f"trima_{p}", lambda: R.calc_trima(c, p))
        attempt(f"wma_{p}", lambda: R.calc_wma(c, p))
    attempt("trima_7", lambda: R.calc_trima(c, 7))
    attempt("bbands_20", lambda: R.bbands(c, 20, 2.0, 2.0))
    attempt("bbands_5", lambda: R.bbands(c, 5, 1.5, 2.5))
    attempt("midpoint_14", lambda: R.midpoint(c, 14))
    attempt("midprice_14", lambda: R.midprice(h, l, 14))
    attempt("rsi_14", lambda: R.rsi(c, 14))
    attempt("rsi_5", lambda: R.rsi(c, 5))
    attempt("macd_12_26_9", lambda: R.macd(c, 12, 26, 9))
    attempt("macd_3_5_8", lambda: R.macd(c, 3, 5, 8))
    attempt("trange", lambda: R.calc_trange(h, l, c))
    attempt("atr_14", lambda: R.atr(h, l, c, 14))
    attempt("natr_14", lambda: R.natr(h, l, c, 14))
    attempt("obv", lambda: R.obv(c, v))
    attempt("ad", lambda: R.calc_ad(h, l, c, v))
    attempt("adosc_3_10", lambda: R.adosc(h, l, c, v, 3, 10))
    attempt("willr_14", lambda: R.willr(h, l, c, 14))
    attempt("stoch_5_3_3", lambda: R.stoch(h, l, c, 5, 3, 0, 3, 0))
    attempt("stochf_5_3", lambda: R.stochf(h, l, c, 5, 3, 0))
    attempt("kdj_9_3_3", lambda: R.kdj(h, l, c, 9, 3, 3))
    attempt("mom_10", lambda: R.mom(c, 10))
    for kind in range(4):
        attempt(f"roc_10_{kind}", lambda: R.roc(c, 10, kind))
    attempt("cmo_14", lambda: R.cmo(c, 14))
    attempt("mfi_14", lambda: R.mfi(h, l, c, v, 14))
    attempt("cci_14", lambda: R.cci(h, l, c, 14))


def main():
    arrays = {}
    # case A: BASELINE config 1 -- 1 symbol x 252 daily bars, dense
    d = synth.ohlcv(1, 252, seed=20260101)
    cols = {k: d[k][0] for k in ("close", "high", "low", "volume")}
    for k, a in cols.items():
        arrays[f"A/in/{k}"] = a
    run_all(*(synth.to_opt(cols[k]) for k in ("close", "high", "low", "volume")), "A", arrays)

    # case B: 300 bars, 7 leading nulls on every field + ~1% interior nulls (independent per field)
    d = synth.ohlcv(1, 300, seed=20260102)
    rng = np.random.default_rng(7)
    ins = {}
    for k in ("close", "high", "low", "volume"):
        ok = rng.random(300) > 0.01
        ok[:7] = False
        arrays[f"B/in/{k}"] = d[k][0]
        arrays[f"B/in/{k}_ok"] = ok
        ins[k] = synth.to_opt(d[k][0], ok)
    run_all(ins["close"], ins["high"], ins["low"], ins["volume"], "B", arrays)

    # case C: leading nulls only (a listing that starts late): 260 bars, first 11 null
    d = synth.ohlcv(1, 260, seed=20260103)
    ok = np.ones(260, bool)
    ok[:11] = False
    ins = {}
    for k in ("close", "high", "low", "volume"):
        arrays[f"C/in/{k}"] = d[k][0]
        arrays[f"C/in/{k}_ok"] = ok
        ins[k] = synth.to_opt(d[k][0], ok)
    run_all(ins["close"], ins["high"], ins["low"], ins["volume"], "C", arrays)

    # case D: degenerate values -- flat stretches (diff == 0 branches), ties, integer prices
    n = 120
    close = np.concatenate([np.full(40, 50.0), 50.0 + np.arange(40) % 3, np.full(40, 48.0)])
    high = close + np.where(np.arange(n) % 7 == 0, 0.0, 1.0)
    low = close - np.where(np.arange(n) % 7 == 0, 0.0, 0.5)
    vol = np.full(n, 1000.0)
    for k, a in (("close", close), ("high", high), ("low", low), ("volume", vol)):
        arrays[f"D/in/{k}"] = a
    run_all(*(synth.to_opt(a) for a in (close, high, low, vol)), "D", arrays)

    # case E: short columns (guards): n = 0, 1, 2, 10 with the default periods
    for n in (0, 1, 2, 10):
        d = synth.ohlcv(1, max(n, 1), seed=20260104 + n)
        cols = {k: d[k][0][:n] for k in ("close", "high", "low", "volume")}
        for k, a in cols.items():
            arrays[f"E{n}/in/{k}"] = a
        run_all(*(synth.to_opt(cols[k]) for k in ("close", "high", "low", "volume")), f"E{n}", arrays)

    np.savez_compressed(OUT / "talib_golden.npz", **arrays)
    print("wrote", OUT / "talib_golden.npz", len(arrays), "arrays")


if __name__ == "__main__":
    main()
