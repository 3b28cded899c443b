"""The polars plugin boundary on the GPU (BASELINE config 1: one symbol x 252 daily bars through the
Python API): every `_polars_plugin_<name>` against the oracle, bit for bit, values and Arrow nulls --
called the way polars calls it (Arrow C Data Interface, literal parameters or pickled kwargs), with
chunked / sliced / non-Float64 inputs, and with the reference's null rules."""
import numpy as np
import pyarrow as pa
import pytest

import synth
import tolerances as T
from oracle import pqo
from polars_quant_b200 import plugin, talib

pytestmark = pytest.mark.gpu


def _np(a: pa.Array):
    ok = ~np.asarray(a.is_null())
    return np.asarray(a.to_numpy(zero_copy_only=False), dtype=np.float64), ok


def _same(name, got: pa.Array, ref):
    gv, gok = _np(got)
    gv = np.where(gok, gv, np.nan)
    nbad, msg = T.compare(name, gv, gok, ref[0], ref[1])
    assert nbad == 0, msg


@pytest.fixture(scope="module")
def d():
    x = synth.ohlcv(1, 252, seed=2024)
    return {k: pa.array(v[0]) for k, v in x.items()}, {k: v[0] for k, v in x.items()}


def test_config1_python_api_matches_the_oracle_bit_for_bit(d):
    a, n = d
    c, h, l, v = n["close"], n["high"], n["low"], n["volume"]
    _same("SMA", talib.SMA(a["close"]), pqo.sma(c, 30))
    _same("EMA", talib.EMA(a["close"], 20), pqo.ema(c, 20))
    _same("RSI", talib.RSI(a["close"]), pqo.rsi(c, 14))
    for got, ref, nm in zip(talib.MACD(a["close"]), pqo.macd(c), ("macd", "macd_signal", "macd_hist")):
        _same(nm, got, ref)
    for got, ref, nm in zip(talib.BBANDS(a["close"], 20, 2.0, 2.0), pqo.bbands(c), ("bb_upper", "bb_middle", "bb_lower")):
        _same(nm, got, ref)
    _same("TEMA", talib.TEMA(a["close"], 10), pqo.tema(c, 10))
    _same("TRIMA", talib.TRIMA(a["close"], 9), pqo.trima(c, 9))
    _same("MA(1)", talib.MA(a["close"], 12, 1), pqo.ma(c, 12, 1))
    _same("MIDPOINT", talib.MIDPOINT(a["close"]), pqo.midpoint(c, 14))
    _same("MIDPRICE", talib.MIDPRICE(a["high"], a["low"], 20), pqo.midprice(h, l, 20))
    _same("TRANGE", talib.TRANGE(a["high"], a["low"], a["close"]), pqo.trange(h, l, c))
    _same("ATR", talib.ATR(a["high"], a["low"], a["close"]), pqo.atr(h, l, c, 14))
    _same("NATR", talib.NATR(a["high"], a["low"], a["close"], 7), pqo.natr(h, l, c, 7))
    _same("OBV", talib.OBV(a["close"], a["volume"]), pqo.obv(c, v))
    _same("AD", talib.AD(a["high"], a["low"], a["close"], a["volume"]), pqo.ad(h, l, c, v))
    _same("ADOSC", talib.ADOSC(a["high"], a["low"], a["close"], a["volume"]), pqo.adosc(h, l, c, v, 3, 10))
    _same("WILLR", talib.WILLR(a["high"], a["low"], a["close"]), pqo.willr(h, l, c, 14))
    _same("MOM", talib.MOM(a["close"]), pqo.mom(c, 10))
    for kind, f in enumerate((talib.ROC, talib.ROCP, talib.ROCR, talib.ROCR100)):
        _same(f.__name__, f(a["close"], 5), pqo.roc(c, 5, kind))
    _same("CMO", talib.CMO(a["close"]), pqo.cmo(c, 14))
    _same("MFI", talib.MFI(a["high"], a["low"], a["close"], a["volume"]), pqo.mfi(h, l, c, v, 14))
    _same("CCI", talib.CCI(a["high"], a["low"], a["close"]), pqo.cci(h, l, c, 14))
    for got, ref, nm in zip(talib.STOCH(a["high"], a["low"], a["close"]), pqo.stoch(h, l, c), ("slowk", "slowd")):
        _same(nm, got, ref)
    for got, ref, nm in zip(talib.KDJ(a["high"], a["low"], a["close"]), pqo.kdj(h, l, c), ("k", "d", "j")):
        _same(nm, got, ref)


def test_kwargs_literals_and_defaults_are_the_same_call(d):
    a, n = d
    ref = pqo.ema(n["close"], 17)
    _same("literal", plugin.call("ema", [a["close"], 17]), ref)
    _same("kwargs", plugin.call("ema", [a["close"]], kwargs={"timeperiod": 17}), ref)
    _same("kwargs win", plugin.call("ema", [a["close"], 5], kwargs={"timeperiod": 17}), ref)
    _same("null literal -> default", plugin.call("ema", [a["close"], None]), pqo.ema(n["close"], 30))
    _same("int64 literal", plugin.call("ema", [a["close"], pa.array([17], type=pa.int64())]), ref)
    out = plugin.call("bbands", [a["close"]], kwargs={"timeperiod": 10, "nbdevup": 1.5, "nbdevdn": 2.5})
    for i, r in enumerate(pqo.bbands(n["close"], 10, 1.5, 2.5)):
        _same("bbands kw %d" % i, out.field(i), r)


def test_chunked_sliced_and_non_float64_inputs(d):
    a, n = d
    c = n["close"]
    ref = pqo.sma(c, 10)
    chunks = pa.chunked_array([a["close"].slice(0, 100), a["close"].slice(100, 1), a["close"].slice(101)])
    _same("3 chunks", plugin.call("sma", [chunks, 10]), ref)             # state carries across chunks, overlap.rs:674
    big = pa.array(np.concatenate([[7.0, 8.0, 9.0], c, [1.0]]))
    _same("offset", plugin.call("sma", [big.slice(3, 252), 10]), ref)    # non-zero Arrow offset
    ci = np.round(c).astype(np.int64)
    _same("int64", plugin.call("sma", [pa.array(ci), 10]), pqo.sma(ci.astype(np.float64), 10))
    cf = c.astype(np.float32)
    _same("float32", plugin.call("ema", [pa.array(cf), 10]), pqo.ema(cf.astype(np.float64), 10))
    vi = n["volume"].astype(np.uint32)
    _same("obv uint32 volume", plugin.call("obv", [a["close"], pa.array(vi)]), pqo.obv(c, vi.astype(np.float64)))


def test_null_rules_follow_the_reference_function_by_function(d):
    a, n = d
    c = n["close"].copy()
    ok = np.ones(252, dtype=bool)
    ok[:7] = False                    # listed later
    ok[[40, 41, 120]] = False         # interior nulls
    col = pa.array(c, mask=~ok)
    _same("sma skips nulls", plugin.call("sma", [col, 10]), pqo.sma(c, 10, ok.astype(np.uint8)))
    _same("ema skips nulls", plugin.call("ema", [col, 10]), pqo.ema(c, 10, ok.astype(np.uint8)))
    for i, r in enumerate(pqo.bbands(c, 20, 2.0, 2.0, ok.astype(np.uint8))):
        _same("bbands nulls %d" % i, plugin.call("bbands", [col, 20, 2.0, 2.0]).field(i), r)
    with pytest.raises(plugin.PluginError, match="not contiguous"):      # cont_slice()? momentum.rs:509
        plugin.call("rsi", [col, 14])
    with pytest.raises(plugin.PluginError, match="not contiguous"):
        plugin.call("macd", [col, 12, 26, 9])
    out = plugin.call("sma", [a["close"], 0])                            # timeperiod == 0 -> all null, overlap.rs:874
    assert out.null_count == 252
    out = plugin.call("sma", [a["close"], 300])                          # len < timeperiod -> all null
    assert out.null_count == 252
    with pytest.raises(plugin.PluginError, match="matype"):
        plugin.call("ma", [a["close"], 10, 2])                           # WMA: defective in the reference, not built


def test_concurrent_calls_from_several_threads(d):
    """polars evaluates independent expressions on its rayon pool: the symbols must be re-entrant."""
    import threading
    a, n = d
    c = n["close"]
    refs = {p: pqo.ema(c, p) for p in range(2, 18)}
    errs = []

    def work(p):
        try:
            for _ in range(5):
                _same("ema %d" % p, plugin.call("ema", [a["close"], p]), refs[p])
        except Exception as e:          # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(p,)) for p in refs]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs[:2]
