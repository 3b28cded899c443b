"""The C oracle (oracle/pq_oracle.c) must reproduce, bit for bit, the golden vectors made by
the independent pure-Python restatement (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import pqo

CASES = ["A", "B", "C", "D", "E0", "E1", "E2", "E10"]


def _inputs(g, tag):
    cols = {}
    for k in ("close", "high", "low", "volume"):
        v = g[f"{tag}/in/{k}"]
        ok = g[f"{tag}/in/{k}_ok"] if f"{tag}/in/{k}_ok" in g.files else None
        cols[k] = (v, ok)
    return cols


def _table(cols):
    (c, cok), (h, hok), (l, lok), (v, vok) = (cols[k] for k in ("close", "high", "low", "volume"))
    t = {}
    for p in (1, 2, 5, 30):
        t[f"sma_{p}"] = lambda p=p: pqo.sma(c, p, cok)
        t[f"ema_{p}"] = lambda p=p: pqo.ema(c, p, cok)
        t[f"tema_{p}"] = lambda p=p: pqo.tema(c, p, cok)
        t[f"trima_{p}"] = lambda p=p: pqo.trima(c, p, cok)
        t[f"wma_{p}"] = lambda p=p: pqo.wma(c, p, cok)
    t["trima_7"] = lambda: pqo.trima(c, 7, cok)
    t["bbands_20"] = lambda: pqo.bbands(c, 20, 2.0, 2.0, cok)
    t["bbands_5"] = lambda: pqo.bbands(c, 5, 1.5, 2.5, cok)
    t["midpoint_14"] = lambda: pqo.midpoint(c, 14, cok)
    t["midprice_14"] = lambda: pqo.midprice(h, l, 14, hok, lok)
    t["rsi_14"] = lambda: pqo.rsi(c, 14, cok)
    t["rsi_5"] = lambda: pqo.rsi(c, 5, cok)
    t["macd_12_26_9"] = lambda: pqo.macd(c, 12, 26, 9, cok)
    t["macd_3_5_8"] = lambda: pqo.macd(c, 3, 5, 8, cok)
    t["trange"] = lambda: pqo.trange(h, l, c, hok, lok, cok)
    t["atr_14"] = lambda: pqo.atr(h, l, c, 14, hok, lok, cok)
    t["natr_14"] = lambda: pqo.natr(h, l, c, 14, hok, lok, cok)
    t["obv"] = lambda: pqo.obv(c, v, cok, vok)
    t["ad"] = lambda: pqo.ad(h, l, c, v, hok, lok, cok, vok)
    t["adosc_3_10"] = lambda: pqo.adosc(h, l, c, v, 3, 10, hok, lok, cok, vok)
    t["willr_14"] = lambda: pqo.willr(h, l, c, 14, hok, lok, cok)
    t["stoch_5_3_3"] = lambda: pqo.stoch(h, l, c, 5, 3, 0, 3, 0, hok, lok, cok)
    t["stochf_5_3"] = lambda: pqo.stochf(h, l, c, 5, 3, 0, hok, lok, cok)
    t["mom_10"] = lambda: pqo.mom(c, 10, cok)
    for kind in range(4):
        t[f"roc_10_{kind}"] = lambda kind=kind: pqo.roc(c, 10, kind, cok)
    t["cmo_14"] = lambda: pqo.cmo(c, 14, cok)
    if cok is None:
        t["kdj_9_3_3"] = lambda: pqo.kdj(h, l, c, 9, 3, 3)
        t["mfi_14"] = lambda: pqo.mfi(h, l, c, v, 14)
    t["cci_14"] = lambda: pqo.cci(h, l, c, 14, hok, lok, cok)
    return t


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("tag", CASES)
def test_c_oracle_matches_golden_bit_exact(golden, tag):
    g = golden
    cols = _inputs(g, tag)
    checked = 0
    for name, fn in _table(cols).items():
        if f"{tag}/{name}/err" in g.files:
            with pytest.raises(pqo.OracleError):
                fn()
            checked += 1
            continue
        res = fn()
        if isinstance(res, tuple) and not isinstance(res[0], tuple):
            res = (res,)
        for k, (vals, ok) in enumerate(res):
            gv, gok = g[f"{tag}/{name}/{k}/v"], g[f"{tag}/{name}/{k}/ok"]
            assert np.array_equal(ok, gok), f"{tag}/{name}/{k}: validity differs"
            assert np.array_equal(_bits(vals[ok]), _bits(gv[gok])), f"{tag}/{name}/{k}: values differ"
            checked += 1
    assert checked > 40


def test_first_valid_index_contract(golden):
    """SURVEY.md 8a 'first non-null index' table, on the dense 1 x 252 case."""
    g = golden
    first = lambda key: int(np.argmax(g[key]))
    assert first("A/sma_30/0/ok") == 29 and first("A/ema_30/0/ok") == 29
    assert first("A/tema_30/0/ok") == 87 and first("A/trima_30/0/ok") == 29
    assert first("A/bbands_20/0/ok") == 19
    assert first("A/macd_12_26_9/0/ok") == 25 and first("A/macd_12_26_9/1/ok") == 8
    assert first("A/macd_12_26_9/2/ok") == 25
    assert np.all(g["A/macd_12_26_9/1/v"][8:25] == 0.0)          # zero-filled signal warm-up
    assert first("A/rsi_14/0/ok") == 13 and first("A/trange/0/ok") == 1
    assert first("A/atr_14/0/ok") == 27 and first("A/natr_14/0/ok") == 27
    assert first("A/obv/0/ok") == 1 and first("A/ad/0/ok") == 0
    assert first("A/adosc_3_10/0/ok") == 9 and first("A/willr_14/0/ok") == 13
    assert first("A/midpoint_14/0/ok") == 0 and first("A/midprice_14/0/ok") == 0
    assert first("A/kdj_9_3_3/0/ok") == 10 and first("A/kdj_9_3_3/1/ok") == 12
    assert first("A/kdj_9_3_3/2/ok") == 12
    assert first("A/stochf_5_3/0/ok") == 4


def test_null_cases_error_like_reference(golden):
    g = golden
    for name in ("rsi_14", "macd_12_26_9", "willr_14", "mom_10", "cmo_14", "cci_14", "midprice_14"):
        assert f"B/{name}/err" in g.files, name
