"""GPU parity: the CUDA path (through the C ABI, libpqb200.so) against the CPU oracle on the same
seeded inputs.  Bar: validity (null positions) bit-exact; OBV bit-exact; floats within
rel 1e-10 / abs 1e-12 (BASELINE.json north_star).  AD is a sign-indefinite running sum whose
reference value carries rounding noise proportional to the largest partial sum so far, so its
tolerance is scaled by the running max |AD| (SURVEY.md section 7 'hard parts')."""
import numpy as np
import pytest

import synth
from oracle import pqo

pytestmark = pytest.mark.gpu

import tolerances as T


@pytest.fixture(scope="module")
def pq():
    import polars_quant_b200 as m
    return m


def _assert_parity(res, out, ok, close, nbdevup=2.0, nbdevdn=2.0, skip=()):
    fails, worst = T.compare_all(res, out, ok, close, pqo.OUTPUT_NAMES, nbdevup, nbdevdn, skip)
    assert not fails, "\n".join(fails)
    return worst


def _run_vs_oracle(pq, d, params=None, oparams=None, starts=None):
    S, N = d["close"].shape
    panel = pq.Panel(S, N)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"], starts=starts)
    res = panel.compute(params)
    if starts is None:
        out, ok, _ = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"], oparams)
    else:
        out = np.full((pqo.N_OUT, S, N), np.nan)
        ok = np.zeros((pqo.N_OUT, S, N), bool)
        for s in range(S):
            a = int(starts[s])
            if a >= N:
                continue
            o, k, _ = pqo.suite_panel(*(d[f][s:s + 1, a:] for f in ("close", "high", "low", "volume")), oparams)
            out[:, s, a:], ok[:, s, a:] = o[:, 0], k[:, 0]
    up = oparams.bb_up if oparams is not None else 2.0
    dn = oparams.bb_dn if oparams is not None else 2.0
    worst = _assert_parity(res, out, ok, d["close"], up, dn)
    panel.close()
    return worst


def test_suite_small_panel(pq):
    d = synth.ohlcv(67, 700, seed=42)            # 5.5 tiles, ragged tail, more symbols than one CTA
    worst = _run_vs_oracle(pq, d)
    print("worst err/tol per output:", {k: f"{v:.2e}" for k, v in worst.items()})


def test_suite_single_symbol_config1(pq):
    d = synth.ohlcv(1, 252, seed=20260101)       # BASELINE config 1 shape
    _run_vs_oracle(pq, d)


@pytest.mark.parametrize("n_bars", [1, 2, 3, 4, 5, 29, 30, 31, 88, 127, 128, 129, 255, 256, 257])
def test_suite_short_and_boundary_lengths(pq, n_bars):
    d = synth.ohlcv(5, n_bars, seed=100 + n_bars)
    _run_vs_oracle(pq, d)


def test_suite_leading_nulls(pq):
    d = synth.ohlcv(40, 600, seed=7)
    rng = np.random.default_rng(1)
    starts = rng.integers(0, 300, size=40).astype(np.int32)
    starts[:6] = [0, 1, 127, 128, 129, 599]
    _run_vs_oracle(pq, d, starts=starts)


@pytest.mark.parametrize("seed", range(8))
def test_suite_random_periods(pq, seed):
    """Random periods.  bbands_period >= 2: with period 1 the reference's sum_sq/p - mean^2 is pure
    rounding drift of its running sums (true variance 0) that sqrt() amplifies to ~1e-6 -- no
    reordered evaluation can reproduce that noise (DESIGN.md 'numerics').  Odd seeds tie
    ema==tema and natr==atr so the steady tile path (stage sharing) runs with non-default periods."""
    from polars_quant_b200 import _native as N
    rng = np.random.default_rng(seed)
    r = lambda lo, hi: int(rng.integers(lo, hi + 1))
    kw = dict(sma_period=r(1, 32), ema_period=r(1, 60), tema_period=r(1, 40), trima_period=r(1, 60),
              bbands_period=r(2, 32), bbands_nbdevup=1.5, bbands_nbdevdn=2.5, macd_fast=r(1, 20),
              macd_slow=r(2, 40), macd_signal=r(1, 15), rsi_period=r(1, 30), atr_period=r(1, 30),
              natr_period=r(1, 30), kdj_fastk=r(1, 32), kdj_slowk=r(1, 8), kdj_slowd=r(1, 8),
              willr_period=r(1, 32), midprice_period=r(1, 32))
    if seed % 2 == 1:
        kw["ema_period"] = kw["tema_period"]
        kw["natr_period"] = kw["atr_period"]
    params = N.default_params(**kw)
    op = pqo.SuiteParams(kw["sma_period"], kw["ema_period"], kw["tema_period"], kw["trima_period"],
                         kw["bbands_period"], 1.5, 2.5, kw["macd_fast"], kw["macd_slow"], kw["macd_signal"],
                         kw["rsi_period"], kw["atr_period"], kw["natr_period"], kw["kdj_fastk"],
                         kw["kdj_slowk"], kw["kdj_slowd"], kw["willr_period"], kw["midprice_period"])
    d = synth.ohlcv(33, 900, seed=900 + seed)
    _run_vs_oracle(pq, d, params, op)


def test_suite_flat_and_tied_values(pq):
    """diff == 0 branches (willr 0, ad literal 0.0, rsi 100) and ties in the rolling extrema."""
    n = 300
    close = np.concatenate([np.full(100, 50.0), 50.0 + np.arange(100) % 3, np.full(100, 48.0)])
    high = close + np.where(np.arange(n) % 7 == 0, 0.0, 1.0)
    low = close - np.where(np.arange(n) % 7 == 0, 0.0, 0.5)
    vol = np.full(n, 1000.0)
    d = {k: np.tile(a, (3, 1)) for k, a in (("close", close), ("high", high), ("low", low), ("volume", vol))}
    S, N = d["close"].shape
    panel = pq.Panel(S, N)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"])
    res = panel.compute()
    out, ok, _ = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"])
    # flat 9-bar windows give 0/0 = NaN in fastk; the reference's running sum then stays NaN forever
    # (sum -= NaN), the tile-local prefix sum recovers once the NaN leaves the window: K/D/J are
    # compared up to the first NaN only (DESIGN.md "NaN inputs").
    _assert_parity(res, out, ok, d["close"], skip=("kdj_k", "kdj_d", "kdj_j"))
    ctx = {"close": d["close"], "out": {n: out[j] for j, n in enumerate(pqo.OUTPUT_NAMES)}}
    for name in ("kdj_k", "kdj_d", "kdj_j"):
        j = pqo.OUTPUT_NAMES.index(name)
        for s in range(S):
            nanpos = np.flatnonzero(np.isnan(out[j, s]) & ok[j, s])
            lim = int(nanpos[0]) if nanpos.size else N
            c1 = {"close": d["close"][s, :lim], "out": {n: v[s, :lim] for n, v in ctx["out"].items()}}
            nbad, _, msg = T.compare(name, res[name][0][s, :lim], res[name][1][s, :lim], out[j, s, :lim], ok[j, s, :lim], c1)
            assert nbad == 0, msg
    panel.close()


def test_config2_full_size_against_oracle(pq):
    """BASELINE config 2: 5,000 x 2,520, full suite, every element against the oracle."""
    S, N = 5000, 2520
    panel = pq.Panel(S, N)
    panel.fill_synthetic(seed=0xC0FFEE, sigma=0.02, to_host=True)
    panel.run()
    panel.download()
    panel.sync()
    res = panel.outputs()
    d = {f: np.ascontiguousarray(panel.host_field(f)[:, :N]) for f in ("close", "high", "low", "volume")}
    out, ok, _ = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"])
    worst = _assert_parity(res, out, ok, d["close"])
    print("config2 worst err/tol per output:", {k: f"{v:.2e}" for k, v in worst.items()})
    panel.close()


def test_host_pipeline_equals_device_path(pq):
    """pqb_suite_run_host (chunked, 3 streams) must give the same bytes as upload/run/download."""
    d = synth.ohlcv(300, 1000, seed=5)
    p1 = pq.Panel(300, 1000)
    p1.set_fields(d["close"], d["high"], d["low"], d["volume"])
    r1 = p1.compute()
    r1 = {k: (v.copy(), o.copy()) for k, (v, o) in r1.items()}
    p2 = pq.Panel(300, 1000)
    p2.set_fields(d["close"], d["high"], d["low"], d["volume"])
    p2.run_host(chunk_symbols=64)
    r2 = p2.outputs()
    for name in r1:
        assert np.array_equal(r1[name][1], r2[name][1]), name
        assert np.array_equal(r1[name][0].view(np.uint64), r2[name][0].view(np.uint64)), name
    p1.close(); p2.close()


def test_properties_at_full_size(pq):
    """Size-independent properties at the config-4 row length (5,040 bars) on a slab of symbols:
    scaling prices by 2 scales SMA/EMA/TEMA/TRIMA/BBANDS/MACD/ATR/MIDPRICE by exactly 2 (power of
    two: bit-exact) and leaves RSI/NATR/WILLR/KDJ unchanged; re-running is idempotent."""
    S, N = 2048, 5040
    p = pq.Panel(S, N)
    p.fill_synthetic(seed=1234, sigma=0.02, to_host=True)
    p.run(); p.download(); p.sync()
    base = {k: (v.copy(), o.copy()) for k, (v, o) in p.outputs().items()}
    p.run(); p.download(); p.sync()
    again = p.outputs()
    for k in base:
        assert np.array_equal(base[k][0].view(np.uint64), again[k][0].view(np.uint64)), f"{k}: not idempotent"
    for f in ("close", "high", "low"):
        p.host_field(f)[:] *= 2.0
    p.upload(); p.run(); p.download(); p.sync()
    scaled = p.outputs()
    for k in ("sma", "ema", "tema", "trima", "bb_upper", "bb_middle", "bb_lower", "macd", "macd_signal",
              "macd_hist", "trange", "atr", "midprice"):
        ok = base[k][1]
        assert np.array_equal(scaled[k][1], ok)
        assert np.array_equal((2.0 * base[k][0][ok]).view(np.uint64), scaled[k][0][ok].view(np.uint64)), k
    for k in ("rsi", "natr", "willr", "kdj_k", "kdj_d", "kdj_j", "obv"):
        ok = base[k][1]
        assert np.array_equal(base[k][0][ok].view(np.uint64), scaled[k][0][ok].view(np.uint64)), k
    p.close()
