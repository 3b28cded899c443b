"""Shared helpers for the tests that read tests/golden/talib_ref_golden.npz -- the vectors made by EXECUTING the
reference's own source text (tests/golden/make_ref_golden.py).  `oracle_call(entry, cols)` maps one golden call to
the C oracle's function of the same name; `gpu_call` (tests/test_gpu_ref_golden.py) does the same for the C ABI."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden" / "talib_ref_golden.npz"

_cache = {}


def load():
    if "g" not in _cache:
        g = np.load(GOLDEN)
        _cache["g"] = g
        _cache["index"] = json.loads(str(g["index"]))
    return _cache["g"], _cache["index"]


def inputs(g, entry):
    """-> {name: (values, validity | None)}; the Cb case presents an all-set bitmap (the reference's
    `Some(bitmap)` branches on dense data)."""
    tag = entry["tag"]
    cols = {}
    for k in ("open", "high", "low", "close", "volume"):
        v = g[f"{tag}/in/{k}"]
        ok = g[f"{tag}/in/{k}_ok"] if f"{tag}/in/{k}_ok" in g.files else None
        if ok is None and entry.get("force_bitmap"):
            ok = np.ones(len(v), dtype=bool)
        cols[k] = (v, ok)
    return cols


def expected(g, entry):
    key = f"{entry['tag']}/{entry['key']}"
    if entry.get("err"):
        return None
    return [(g[f"{key}/{j}/v"], g[f"{key}/{j}/ok"]) for j in range(entry["n_out"])]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same(vals, ok, gv, gok):
    """validity identical; valid slots bit-identical (any NaN == any NaN)."""
    if not np.array_equal(np.asarray(ok, bool), np.asarray(gok, bool)):
        return "validity differs at %s" % np.argwhere(np.asarray(ok, bool) != np.asarray(gok, bool))[:4].ravel().tolist()
    a, b = np.asarray(vals, np.float64)[gok], np.asarray(gv, np.float64)[gok]
    good = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    if not good.all():
        i = int(np.argwhere(~good)[0][0])
        return f"{int((~good).sum())} values differ; first: got {a[i]!r} want {b[i]!r}"
    return ""


def oracle_call(entry, cols):
    """Returns a list of (values, ok) from the C oracle for one golden call, or None if the oracle has no entry
    point for it (the caller counts those).  Raises pqo.OracleError where the oracle says the reference fails."""
    from oracle import pqo

    fn, kw, pr = entry["fn"], entry.get("kwargs", {}), entry["params"]
    (o, ook), (h, hok), (l, lok), (c, cok), (v, vok) = (cols[k] for k in ("open", "high", "low", "close", "volume"))
    one = lambda r: [r]
    if entry["kind"] == "py":
        if fn == "STOCH":
            return list(pqo.stoch(h, l, c, *pr, hok=hok, lok=lok, cok=cok))
        if fn == "STOCHF":
            return list(pqo.stochf(h, l, c, *pr, hok=hok, lok=lok, cok=cok))
        if fn == "STOCHRSI":
            return list(pqo.stochrsi(c, *pr, ok=cok))
        if fn == "MACDEXT":
            return list(pqo.macdext(c, *pr, ok=cok))
        if fn == "MACDFIX":
            return list(pqo.macd(c, 12, 26, pr[0], cok))
        return None
    if fn in ("sma", "ema", "tema", "trima", "wma", "dema", "kama"):
        return one(getattr(pqo, fn)(c, kw.get("timeperiod", 30), cok))
    if fn == "t3":
        return one(pqo.t3(c, kw.get("timeperiod", 5), kw.get("vfactor", 0.0), cok))
    if fn == "ma":
        return one(pqo.ma(c, kw.get("timeperiod", 30), kw.get("matype", 0), cok))
    if fn == "bbands":
        return list(pqo.bbands(c, kw.get("timeperiod", 20), kw.get("nbdevup", 2.0), kw.get("nbdevdn", 2.0), cok))
    if fn == "midpoint":
        return one(pqo.midpoint(c, kw.get("timeperiod", 14), cok))
    if fn == "midprice":
        return one(pqo.midprice(h, l, kw.get("timeperiod", 14), hok, lok))
    if fn in ("atr", "natr"):
        return one(getattr(pqo, fn)(h, l, c, kw.get("timeperiod", 14), hok, lok, cok))
    if fn == "trange":
        return one(pqo.trange(h, l, c, hok, lok, cok))
    if fn == "obv":
        return one(pqo.obv(c, v, cok, vok))
    if fn == "ad":
        return one(pqo.ad(h, l, c, v, hok, lok, cok, vok))
    if fn == "adosc":
        return one(pqo.adosc(h, l, c, v, kw.get("fastperiod", 3), kw.get("slowperiod", 10), hok, lok, cok, vok))
    if fn == "willr":
        return one(pqo.willr(h, l, c, pr[0], hok, lok, cok))
    if fn == "rsi":
        return one(pqo.rsi(c, pr[0], cok))
    if fn == "macd":
        return list(pqo.macd(c, pr[0], pr[1], pr[2], cok))
    if fn == "mom":
        return one(pqo.mom(c, pr[0], cok))
    if fn in ("roc", "rocp", "rocr", "rocr100"):
        return one(pqo.roc(c, pr[0], ("roc", "rocp", "rocr", "rocr100").index(fn), cok))
    if fn == "cmo":
        return one(pqo.cmo(c, pr[0], cok))
    if fn == "cci":
        return one(pqo.cci(h, l, c, pr[0], hok, lok, cok))
    if fn == "trix":
        return one(pqo.trix(c, pr[0], cok))
    nulls = any(x is not None and not np.all(x) for x in (ook, hok, lok, cok, vok))
    if fn in ("mfi", "ultosc", "aroon", "adx", "adxr", "dx", "minus_di", "plus_di", "minus_dm", "plus_dm", "bop"):
        if nulls:                              # these oracle entry points take null-free columns; the reference
            raise pqo.OracleError(-1)          # fails with cont_slice()? on anything else
    if fn == "mfi":
        return one(pqo.mfi(h, l, c, v, pr[0]))
    if fn == "ultosc":
        return one(pqo.ultosc(h, l, c, *pr))
    if fn == "aroon":
        return list(pqo.aroon(h, l, pr[0]))
    if fn in ("adx", "adxr", "dx", "minus_di", "plus_di", "minus_dm", "plus_dm"):
        d = pqo.dm(h, l, c, pr[0])
        return one(d["dx" if fn == "plus_di" else fn])       # momentum.rs:409: plus_di returns calc_dm().0 == DX
    if fn in ("avgprice", "medprice", "typprice", "wclprice", "bop"):
        if nulls:
            return None                        # null-propagating price transforms: covered by the candle tests
        return one((pqo.price(("avgprice", "medprice", "typprice", "wclprice", "bop").index(fn), o, h, l, c),
                    np.ones(len(o), bool)))
    return None
