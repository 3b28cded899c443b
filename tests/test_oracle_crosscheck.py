"""C oracle vs the pure-Python restatement on fresh random inputs (bit-exact), incl. random
null patterns, random periods and Arrow-chunking invariance (state carries across chunks in
the reference, overlap.rs:674, so a column's result cannot depend on how it is chunked: both
restatements take the concatenated column).  CPU only, small sizes."""
import numpy as np
import pytest

from oracle import pqo, ref_py as R
import synth


def _eq(c_res, py_col, what):
    vals, ok = c_res
    pv, pok = synth.from_opt(py_col)
    assert np.array_equal(ok, pok), what + ": validity"
    assert np.array_equal(vals[ok].view(np.uint64), pv[pok].view(np.uint64)), what + ": values"


@pytest.mark.parametrize("seed", range(6))
def test_random_columns(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(40, 160))
    d = synth.ohlcv(1, n, seed=1000 + seed)
    c, h, l, v = (d[k][0] for k in ("close", "high", "low", "volume"))
    ok = rng.random(n) > (0.0 if seed % 2 == 0 else 0.05)
    okp = None if ok.all() else ok
    co, ho, lo, vo = (synth.to_opt(a, okp) for a in (c, h, l, v))
    p = int(rng.integers(1, 25))
    _eq(pqo.sma(c, p, okp), R.calc_sma(co, p), "sma")
    _eq(pqo.ema(c, p, okp), R.calc_ema(co, p), "ema")
    _eq(pqo.tema(c, p, okp), R.calc_tema(co, p), "tema")
    _eq(pqo.trima(c, p, okp), R.calc_trima(co, p), "trima")
    for a, b in zip(pqo.bbands(c, p, 2.0, 2.0, okp), R.bbands(co, p, 2.0, 2.0)):
        _eq(a, b, "bbands")
    _eq(pqo.midpoint(c, p, okp), R.midpoint(co, p), "midpoint")
    _eq(pqo.trange(h, l, c, okp, okp, okp), R.calc_trange(ho, lo, co), "trange")
    _eq(pqo.atr(h, l, c, p, okp, okp, okp), R.atr(ho, lo, co, p), "atr")
    _eq(pqo.natr(h, l, c, p, okp, okp, okp), R.natr(ho, lo, co, p), "natr")
    if okp is None:                                            # calc_dm family: cont_slice()? refuses nulls
        dm_c, dm_p = pqo.dm(h, l, c, p), R.dm_family(ho, lo, co, p)
        for k in ("plus_dm", "minus_dm", "dx", "minus_di", "adx", "adxr"):
            _eq(dm_c[k], dm_p[k], "dm." + k)
        _eq(dm_c["dx"], dm_p["plus_di"], "plus_di == dx (momentum.rs:409)")
        _eq(pqo.trix(c, p), R.trix(co, p), "trix")
        _eq(pqo.ultosc(h, l, c, p, p + 3, 2 * p + 1), R.ultosc(ho, lo, co, p, p + 3, 2 * p + 1), "ultosc")
        for a, b in zip(pqo.aroon(h, l, p), R.aroon(ho, lo, p)):
            _eq(a, b, "aroon")
    _eq(pqo.obv(c, v, okp, okp), R.obv(co, vo), "obv")
    _eq(pqo.ad(h, l, c, v, okp, okp, okp, okp), R.calc_ad(ho, lo, co, vo), "ad")
    _eq(pqo.adosc(h, l, c, v, 3, 10, okp, okp, okp, okp), R.adosc(ho, lo, co, vo, 3, 10), "adosc")
    for a, b in zip(pqo.stoch(h, l, c, p, 3, 0, 3, 0, okp, okp, okp), R.stoch(ho, lo, co, p, 3, 0, 3, 0)):
        _eq(a, b, "stoch")
    if okp is None:
        _eq(pqo.rsi(c, p), R.rsi(co, p), "rsi")
        for a, b in zip(pqo.macd(c, p, p + 7, 5), R.macd(co, p, p + 7, 5)):
            _eq(a, b, "macd")
        _eq(pqo.willr(h, l, c, p), R.willr(ho, lo, co, p), "willr")
        _eq(pqo.midprice(h, l, p), R.midprice(ho, lo, p), "midprice")
        for a, b in zip(pqo.kdj(h, l, c, p, 3, 3), R.kdj(ho, lo, co, p, 3, 3)):
            _eq(a, b, "kdj")
        _eq(pqo.cmo(c, p), R.cmo(co, p), "cmo")
        _eq(pqo.mfi(h, l, c, v, p), R.mfi(ho, lo, co, vo, p), "mfi")
        _eq(pqo.cci(h, l, c, p), R.cci(ho, lo, co, p), "cci")
    else:
        with pytest.raises(pqo.OracleError):
            pqo.rsi(c, p, okp)


def test_rma_d1_matches_ema_structure():
    x = np.abs(np.random.default_rng(3).normal(size=64))
    vals, ok = pqo.rma(x, 14)
    assert int(np.argmax(ok)) == 13
    assert vals[13] == np.add.reduce(x[:14].tolist()) / 14 or np.isclose(vals[13], x[:14].sum() / 14, rtol=1e-15)
    _eq((vals, ok), R.calc_rma([float(t) for t in x], 14), "rma")


def test_midpoint_is_rollmax_plus_cummin():
    """The literal midpoint (overlap.rs:227,264 defect) equals (rollmax_p + cummin)/2."""
    c = synth.ohlcv(1, 200, seed=5)["close"][0]
    vals, ok = pqo.midpoint(c, 14)
    rollmax = np.array([c[max(0, i - 13):i + 1].max() for i in range(200)])
    assert ok.all() and np.array_equal(vals, (rollmax + np.minimum.accumulate(c)) / 2.0)


def test_suite_panel_equals_single_calls():
    d = synth.ohlcv(3, 300, seed=11)
    out, ok, threads = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"], threads=2)
    assert threads >= 1
    s = 1
    c, h, l, v = (d[k][s] for k in ("close", "high", "low", "volume"))
    single = [pqo.sma(c, 30), pqo.ema(c, 30), pqo.tema(c, 30), pqo.trima(c, 30), *pqo.bbands(c, 20),
              *pqo.macd(c), pqo.rsi(c, 14), pqo.trange(h, l, c), pqo.atr(h, l, c, 14),
              pqo.natr(h, l, c, 14), pqo.obv(c, v), pqo.ad(h, l, c, v), *pqo.kdj(h, l, c, 9, 3, 3),
              pqo.willr(h, l, c, 14), pqo.midprice(h, l, 14)]
    assert len(single) == pqo.N_OUT
    for j, (vals, okj) in enumerate(single):
        assert np.array_equal(ok[j, s], okj), pqo.OUTPUT_NAMES[j]
        assert np.array_equal(out[j, s][okj].view(np.uint64), vals[okj].view(np.uint64)), pqo.OUTPUT_NAMES[j]


def test_info_oracle_against_numpy():
    """pqo_info (the definition of the last-row reductions, README.md:832-851) against an independent numpy reading."""
    import synth
    d = synth.ohlcv(6, 120, seed=9)
    for s in range(6):
        c, h, l, v = (d[k][s] for k in ("close", "high", "low", "volume"))
        out, ok = pqo.info(c, h, l, v)
        assert ok.all()
        r = c[-20:] / c[-21:-1] - 1.0
        want = [c[-1], h[-1], l[-1], v[-1], (c[-1] / c[-2] - 1) * 100, (c[-1] / c[-6] - 1) * 100, (c[-1] / c[-21] - 1) * 100,
                r.std(ddof=1) * np.sqrt(252.0) * 100, c[-5:].mean(), c[-10:].mean(), c[-20:].mean(), v[-1] / v[-5:].mean(),
                (h[-1] - l[-1]) / c[-1] * 100]
        assert np.allclose(out, want, rtol=1e-12, atol=0)
        assert out[0] == want[0] and out[4] == want[4] and out[12] == want[12]          # single-operation columns: exact
    out, ok = pqo.info(c, h, l, v, start=120 - 7)                                       # 7 valid rows at the end
    assert ok.tolist() == [True] * 4 + [True, True, False, False, True, False, False, True, True]
