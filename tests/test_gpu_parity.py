"""GPU parity: the CUDA path (through the C ABI, libpqb200.so) against the CPU oracle on the same
seeded inputs.  The kernel evaluates every symbol serially in the reference's own operation order,
so the bar asserted here is BIT-EXACT values and validity for all 21 outputs (tests/tolerances.py;
BASELINE.json asks for validity/OBV bit-exact and floats within rel 1e-10 / abs 1e-12)."""
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


def _assert_parity(res, out, ok, skip=()):
    fails = T.compare_all(res, out, ok, pqo.OUTPUT_NAMES, skip)
    assert not fails, "\n".join(fails)


def _oracle(d, oparams=None, starts=None):
    S, N = d["close"].shape
    if starts is None:
        out, ok, _ = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"], oparams)
        return out, ok
    out = np.full((pqo.N_OUT, S, N), np.nan)
    ok = np.zeros((pqo.N_OUT, S, N), bool)
    for s in range(S):
        a = int(starts[s])
        if a >= N:
            continue
        o, k, _ = pqo.suite_panel(*(d[f][s:s + 1, a:] for f in ("close", "high", "low", "volume")), oparams)
        out[:, s, a:], ok[:, s, a:] = o[:, 0], k[:, 0]
    return out, ok


def _run_vs_oracle(pq, d, params=None, oparams=None, starts=None):
    S, N = d["close"].shape
    panel = pq.Panel(S, N)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"], starts=starts)
    res = panel.compute(params)
    out, ok = _oracle(d, oparams, starts)
    _assert_parity(res, out, ok)
    panel.close()


def test_suite_small_panel(pq):
    d = synth.ohlcv(67, 700, seed=42)            # 3 symbol blocks (ragged last one), 88 stages, ragged tail
    _run_vs_oracle(pq, d)


def test_suite_single_symbol_config1(pq):
    d = synth.ohlcv(1, 252, seed=20260101)       # BASELINE config 1 shape
    _run_vs_oracle(pq, d)


@pytest.mark.parametrize("n_bars", [1, 2, 3, 4, 5, 7, 8, 9, 29, 30, 31, 88, 127, 128, 129, 255, 256, 257])
def test_suite_short_and_boundary_lengths(pq, n_bars):
    d = synth.ohlcv(5, n_bars, seed=100 + n_bars)
    _run_vs_oracle(pq, d)


@pytest.mark.parametrize("n_symbols", [1, 31, 32, 33, 64, 65])
def test_suite_block_boundaries(pq, n_symbols):
    d = synth.ohlcv(n_symbols, 300, seed=500 + n_symbols)
    _run_vs_oracle(pq, d)


@pytest.mark.parametrize("n_symbols", [4736, 5100, 5300, 5500, 6272, 7850, 7900, 9472])
def test_suite_around_one_block_per_sm(pq, n_symbols):
    """148 SMs: 148 blocks (nine-warp small-panel variant), 160 / 166 / 172 / 196 (one GPU's share of config 4 at 8 GPUs) /
    246 blocks (the same + the blocks beyond one per SM as compact tail CTAs on a second stream: three role warps + a producer
    each), 247 and 296 blocks (the plain kernel) -- every launch shape of launch_suite against the oracle, all symbols."""
    d = synth.ohlcv(n_symbols, 200, seed=900 + n_symbols)
    _run_vs_oracle(pq, d)


def test_suite_leading_nulls(pq):
    """Symbols listed at different dates: per-symbol first valid bar (leading Arrow nulls)."""
    d = synth.ohlcv(40, 600, seed=7)
    rng = np.random.default_rng(1)
    starts = rng.integers(0, 300, size=40).astype(np.int32)
    starts[:8] = [0, 1, 7, 8, 9, 127, 599, 600]
    _run_vs_oracle(pq, d, starts=starts)


def _random_params(seed):
    from polars_quant_b200 import _native as N
    rng = np.random.default_rng(seed)
    r = lambda lo, hi: int(rng.integers(lo, hi + 1))
    kw = dict(sma_period=r(1, 70), ema_period=r(1, 60), tema_period=r(1, 40), trima_period=r(1, 60),
              bbands_period=r(1, 50), bbands_nbdevup=1.5, bbands_nbdevdn=2.5, macd_fast=r(1, 20),
              macd_slow=r(2, 40), macd_signal=r(1, 15), rsi_period=r(1, 30), atr_period=r(1, 30),
              natr_period=r(1, 30), kdj_fastk=r(1, 40), kdj_slowk=r(1, 8), kdj_slowd=r(1, 8),
              willr_period=r(1, 40), midprice_period=r(1, 40))
    if seed % 3 == 1:
        kw["ema_period"] = kw["tema_period"]
        kw["natr_period"] = kw["atr_period"]
        kw["midprice_period"] = kw["willr_period"]
    params = N.default_params(**kw)
    op = pqo.SuiteParams(kw["sma_period"], kw["ema_period"], kw["tema_period"], kw["trima_period"],
                         kw["bbands_period"], 1.5, 2.5, kw["macd_fast"], kw["macd_slow"], kw["macd_signal"],
                         kw["rsi_period"], kw["atr_period"], kw["natr_period"], kw["kdj_fastk"],
                         kw["kdj_slowk"], kw["kdj_slowd"], kw["willr_period"], kw["midprice_period"])
    return params, op


@pytest.mark.parametrize("seed", range(12))
def test_suite_random_periods(pq, seed):
    """Random periods (period 1 included: its quirks -- TEMA(1), BBANDS(1) running-sum noise under
    sqrt -- are reproduced exactly because the arithmetic is the reference's own)."""
    params, op = _random_params(seed)
    d = synth.ohlcv(33, 900, seed=900 + seed)
    _run_vs_oracle(pq, d, params, op)


def test_suite_flat_and_tied_values(pq):
    """diff == 0 branches (willr 0, ad literal 0.0, rsi 100), ties in the rolling extrema, and
    flat k-bar windows whose fastk is 0/0 = NaN: the reference's running sums stay NaN forever
    afterwards (sum -= NaN); the serial kernel reproduces exactly that."""
    n = 300
    close = np.concatenate([np.full(100, 50.0), 50.0 + np.arange(100) % 3, np.full(100, 48.0)])
    high = close + np.where(np.arange(n) % 7 == 0, 0.0, 1.0)
    low = close - np.where(np.arange(n) % 7 == 0, 0.0, 0.5)
    vol = np.full(n, 1000.0)
    d = {k: np.tile(a, (3, 1)) for k, a in (("close", close), ("high", high), ("low", low), ("volume", vol))}
    d["high"][1] = d["close"][1]                 # symbol 1: high == low == close on every bar
    d["low"][1] = d["close"][1]
    d["volume"][2, 50:80] = 0.0                  # symbol 2: halted (zero volume)
    _run_vs_oracle(pq, d)


def test_nan_and_inf_inputs_propagate_like_the_reference(pq):
    """A NaN *value* is an ordinary number to the reference: it poisons running sums and EMAs for
    the rest of the column.  Close-only indicators and OBV/AD/TRANGE/ATR must match
    the oracle bit for bit; rolling max/min with NaN operands are order-dependent in the reference
    (deque vs brute force) and are left out."""
    d = synth.ohlcv(6, 400, seed=77)
    d["close"][0, 150] = np.nan
    d["close"][1, 10] = np.inf
    d["volume"][2, 33] = np.nan
    d["close"][3, 0] = np.nan
    S, N = d["close"].shape
    panel = pq.Panel(S, N)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"])
    res = panel.compute()
    out, ok = _oracle(d)
    _assert_parity(res, out, ok, skip=("kdj_k", "kdj_d", "kdj_j", "willr", "midprice"))
    panel.close()


def test_nan_values_in_high_and_low_willr_follows_f64_max(pq):
    """willr folds its window with f64::max / f64::min, which ignore a NaN operand (momentum.rs:644-650); the reference
    is deterministic there and the CUDA path's running extremes follow the same rule -- bit-exact, also through the
    full-suite kernel where WILLR and MIDPRICE share one pair of van Herk arrays.  (midprice / midpoint keep monotonic
    deques whose compares are false on NaN, overlap.rs:206-221: while a NaN sits in such a window the reference returns
    the maximum of the bars BEFORE it -- an artifact of the container the CUDA path does not reproduce: DESIGN.md 5.)"""
    d = synth.ohlcv(6, 400, seed=78)
    d["high"][0, 90] = np.nan
    d["low"][1, 130] = np.nan
    d["high"][2, 131] = np.nan
    d["low"][2, 131] = np.nan
    d["high"][3, 5:25] = np.nan                  # a whole window of NaN highs (p = 14)
    S, N = d["close"].shape
    panel = pq.Panel(S, N)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"])
    res = panel.compute()
    for s in range(S):
        rv, rk = pqo.willr(d["high"][s], d["low"][s], d["close"][s], 14)
        nbad, msg = T.compare("willr", res["willr"][0][s], res["willr"][1][s], rv, rk)
        assert nbad == 0, f"symbol {s}: {msg}"
    # outside the windows that hold a NaN everything else is untouched
    out, ok = _oracle(d)
    clean = np.ones((S, N), bool)
    for s, t in ((0, 90), (1, 130), (2, 131)):
        clean[s, t:t + 14] = False
    clean[3, 5:25 + 14] = False
    for name in ("midprice",):
        j = pqo.OUTPUT_NAMES.index(name)
        assert T.same_bits(res[name][0][clean], out[j][clean]).all(), name
    panel.close()


def test_partial_suites_and_single_indicator_masks(pq):
    """Any subset of indicator groups (the non-specialised kernel): outputs of enabled groups match,
    disabled outputs stay unallocated."""
    from polars_quant_b200 import _native as N
    d = synth.ohlcv(35, 500, seed=321)
    out, ok = _oracle(d)
    rng = np.random.default_rng(5)
    masks = [N.IND[k] for k in N.IND] + [int(rng.integers(1, N.IND_ALL)) for _ in range(6)]
    groups_of = {"sma": ["sma"], "ema": ["ema"], "tema": ["tema"], "trima": ["trima"],
                 "bbands": ["bb_upper", "bb_middle", "bb_lower"], "macd": ["macd", "macd_signal", "macd_hist"],
                 "rsi": ["rsi"], "trange": ["trange"], "atr": ["atr"], "natr": ["natr"], "obv": ["obv"], "ad": ["ad"],
                 "kdj": ["kdj_k", "kdj_d", "kdj_j"], "willr": ["willr"], "midprice": ["midprice"]}
    for m in masks:
        names = [o for g, outs in groups_of.items() if m & N.IND[g] for o in outs]
        omask = sum(1 << pqo.OUTPUT_NAMES.index(o) for o in names)
        panel = pq.Panel(35, 500, outputs_mask=omask)
        panel.set_fields(d["close"], d["high"], d["low"], d["volume"])
        res = panel.compute(N.default_params(indicators=m))
        assert sorted(res) == sorted(names)
        _assert_parity(res, out, ok)
        panel.close()


def test_partial_suites_on_a_many_wave_panel_narrow_ctas(pq):
    """Beyond three blocks per SM a partial suite is launched with fewer than eight warps per CTA (its slots + the producer at
    least) and TMA stages that hold only the fields it reads, so that more blocks are resident (engine.cu launch_suite): one,
    two, three and four slots, close-only / high-low-close / every field, 16,640 symbols = 520 blocks, every symbol against
    the oracle."""
    from polars_quant_b200 import _native as N
    S, NB = 16_640, 160
    d = synth.ohlcv(S, NB, seed=77)
    starts = np.zeros(S, np.int32)
    starts[::97] = 40                                         # a few late listings
    out, ok = None, None
    groups_of = {"ema": ["ema"], "rsi": ["rsi"], "bbands": ["bb_upper", "bb_middle", "bb_lower"], "kdj": ["kdj_k", "kdj_d", "kdj_j"],
                 "atr": ["atr"], "obv": ["obv"], "ad": ["ad"], "sma": ["sma"], "macd": ["macd", "macd_signal", "macd_hist"],
                 "willr": ["willr"], "midprice": ["midprice"]}
    for groups in (["ema"], ["rsi"], ["bbands"], ["kdj", "atr"], ["obv", "ad"], ["sma", "ema", "rsi", "macd", "bbands"], ["willr", "midprice"]):
        if out is None:
            out, ok = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"], None)[:2]
            for s in np.nonzero(starts)[0]:
                o, k, _ = pqo.suite_panel(*(d[f][s:s + 1, 40:] for f in ("close", "high", "low", "volume")), None)
                out[:, s, :40], ok[:, s, :40] = np.nan, False
                out[:, s, 40:], ok[:, s, 40:] = o[:, 0], k[:, 0]
        names = [o for g in groups for o in groups_of[g]]
        panel = pq.Panel(S, NB, outputs_mask=sum(1 << pqo.OUTPUT_NAMES.index(o) for o in names))
        panel.set_fields(d["close"], d["high"], d["low"], d["volume"], starts=starts)
        res = panel.compute(N.default_params(indicators=sum(N.IND[g] for g in groups)))
        assert sorted(res) == sorted(names)
        _assert_parity(res, out, ok)
        panel.close()


def test_period_zero_gives_all_null_columns(pq):
    """timeperiod == 0 -> the reference's guards return an all-null column (overlap.rs:663,874)."""
    from polars_quant_b200 import _native as N
    d = synth.ohlcv(3, 100, seed=2)
    panel = pq.Panel(3, 100)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"])
    res = panel.compute(N.default_params(sma_period=0, ema_period=0, rsi_period=0, willr_period=0))
    for name in ("sma", "ema", "rsi", "willr"):
        v, k = res[name]
        assert not k.any() and np.isnan(v).all(), name
    out, ok = _oracle(d)
    _assert_parity(res, out, ok, skip=("sma", "ema", "rsi", "willr"))
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
    out, ok = _oracle(d)
    _assert_parity(res, out, ok)
    panel.close()


def test_config4_row_length_slab_against_oracle(pq):
    """BASELINE config 4 row length (5,040 bars): a slab of 1,000 symbols against the oracle."""
    S, N = 1000, 5040
    panel = pq.Panel(S, N)
    panel.fill_synthetic(seed=4, sigma=0.02, to_host=True)
    panel.run_host()
    res = panel.outputs()
    d = {f: np.ascontiguousarray(panel.host_field(f)[:, :N]) for f in ("close", "high", "low", "volume")}
    out, ok = _oracle(d)
    _assert_parity(res, out, ok)
    panel.close()


def test_config3_long_rows_ema_macd(pq):
    """BASELINE config 3 (scan-depth stress): 1,000,000 minute bars per row, EMA(12/26/200/5000) and
    MACD(12,26,9), against the oracle's serial loops.  64 symbols keep the oracle and the pinned
    staging small; the row length is the config's."""
    from polars_quant_b200 import _native as N
    S, NB = 64, 1_000_000
    omask = sum(1 << pqo.OUTPUT_NAMES.index(o) for o in ("ema", "macd", "macd_signal", "macd_hist"))
    panel = pq.Panel(S, NB, fields_mask=1, outputs_mask=omask)
    d = synth.ohlcv(S, NB, seed=3, sigma=0.0005)
    panel.set_fields(close=d["close"])
    for period in (12, 26, 200, 5000):
        res = panel.compute(N.default_params(indicators=N.IND["ema"] | N.IND["macd"], ema_period=period))
        for s in range(0, S, 9):
            v, k = pqo.ema(d["close"][s], period)
            nbad, msg = T.compare("ema", res["ema"][0][s], res["ema"][1][s], v, k)
            assert nbad == 0, f"period {period} symbol {s}: {msg}"
    for s in range(0, S, 9):
        ref = pqo.macd(d["close"][s], 12, 26, 9)
        for name, (v, k) in zip(("macd", "macd_signal", "macd_hist"), ref):
            nbad, msg = T.compare(name, res[name][0][s], res[name][1][s], v, k)
            assert nbad == 0, f"symbol {s}: {msg}"
    panel.close()


def test_config5_long_windows_kdj_donchian_atr(pq):
    """BASELINE config 5: rolling max/min windows 5..250 (KDJ fastk, WILLR, MIDPRICE = Donchian mid)
    plus ATR(14) on 5,040-bar rows; one launch per window (the per-block van Herk arrays of two
    250-bar windows do not fit shared memory together)."""
    from polars_quant_b200 import _native as N
    S, NB = 256, 5040
    d = synth.ohlcv(S, NB, seed=55)
    names = ("atr", "kdj_k", "kdj_d", "kdj_j", "willr", "midprice")
    omask = sum(1 << pqo.OUTPUT_NAMES.index(o) for o in names)
    omask |= sum(1 << N.OUTPUT_NAMES.index(o) for o in ("donchian_upper", "donchian_lower"))
    panel = pq.Panel(S, NB, outputs_mask=omask)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"])
    pick = list(range(0, S, 37))
    for k in (5, 9, 14, 60, 250):
        res = panel.compute(N.default_params(indicators=N.IND["kdj"] | N.IND["atr"], kdj_fastk=k))
        for s in pick:
            ref = pqo.kdj(d["high"][s], d["low"][s], d["close"][s], k, 3, 3)
            for name, (v, kk) in zip(("kdj_k", "kdj_d", "kdj_j"), ref):
                nbad, msg = T.compare(name, res[name][0][s], res[name][1][s], v, kk)
                assert nbad == 0, f"fastk {k} symbol {s}: {msg}"
            v, kk = pqo.atr(d["high"][s], d["low"][s], d["close"][s], 14)
            nbad, msg = T.compare("atr", res["atr"][0][s], res["atr"][1][s], v, kk)
            assert nbad == 0, msg
    for p in (5, 20, 55, 250):
        res = panel.compute(N.default_params(indicators=N.IND["willr"] | N.IND["midprice"] | N.IND_EXTRA["donchian"],
                                             willr_period=p, midprice_period=p, donchian_period=p))
        for s in pick:
            v, kk = pqo.willr(d["high"][s], d["low"][s], d["close"][s], p)
            nbad, msg = T.compare("willr", res["willr"][0][s], res["willr"][1][s], v, kk)
            assert nbad == 0, f"willr {p} symbol {s}: {msg}"
            v, kk = pqo.midprice(d["high"][s], d["low"][s], p)
            nbad, msg = T.compare("midprice", res["midprice"][0][s], res["midprice"][1][s], v, kk)
            assert nbad == 0, f"midprice {p} symbol {s}: {msg}"
            up, lo = pqo.donchian(d["high"][s], d["low"][s], p)          # Donchian channel (D3); its mid == midprice
            assert T.same_bits((up[0] + lo[0]) / 2.0, res["midprice"][0][s]).all()
            for name, (v, kk) in (("donchian_upper", up), ("donchian_lower", lo)):
                nbad, msg = T.compare(name, res[name][0][s], res[name][1][s], v, kk)
                assert nbad == 0, f"{name} {p} symbol {s}: {msg}"
    # two different 250-bar windows in one launch exceed the shared-memory budget: loud, not silent
    with pytest.raises(N.PqbError) as ei:
        panel.compute(N.default_params(indicators=N.IND["willr"] | N.IND["midprice"] | N.IND["kdj"], willr_period=250,
                                       midprice_period=249, kdj_fastk=250))
    assert ei.value.code == -4
    panel.close()


def test_host_pipeline_equals_device_path(pq):
    """pqb_suite_run_host (chunked, 3 streams, pack/unpack) must give the same bytes as upload/run/download."""
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
        assert T.same_bits(r1[name][0], r2[name][0]).all(), name
    p1.close(); p2.close()


def test_properties_at_full_size(pq):
    """Size-independent properties at the config-4 row length (5,040 bars) on a slab of symbols:
    scaling prices by 2 scales SMA/EMA/TEMA/TRIMA/BBANDS/MACD/ATR/MIDPRICE by exactly 2 (power of
    two: bit-exact) and leaves RSI/NATR/WILLR/KDJ/OBV unchanged; re-running is idempotent; every
    symbol of the panel equals the same symbol computed alone (symbols are independent)."""
    S, N = 2048, 5040
    p = pq.Panel(S, N)
    p.fill_synthetic(seed=1234, sigma=0.02, to_host=True)
    p.run(); p.download(); p.sync()
    base = {k: (v.copy(), o.copy()) for k, (v, o) in p.outputs().items()}
    p.run(); p.download(); p.sync()
    again = p.outputs()
    for k in base:
        assert T.same_bits(base[k][0], again[k][0]).all(), f"{k}: not idempotent"
    one = pq.Panel(1, N)
    for s in (0, 31, 32, 1000, 2047):
        one.set_fields(*(p.host_field(f)[s:s + 1, :N] for f in ("close", "high", "low", "volume")))
        r1 = one.compute()
        for k in base:
            assert T.same_bits(base[k][0][s], r1[k][0][0]).all(), f"{k}: symbol {s} differs when computed alone"
    one.close()
    for f in ("close", "high", "low"):
        p.host_field(f)[:] *= 2.0
    p.upload(); p.run(); p.download(); p.sync()
    scaled = p.outputs()
    for k in ("sma", "ema", "tema", "trima", "bb_upper", "bb_middle", "bb_lower", "macd", "macd_signal",
              "macd_hist", "trange", "atr", "midprice"):
        ok = base[k][1]
        assert np.array_equal(scaled[k][1], ok)
        assert T.same_bits(2.0 * base[k][0][ok], scaled[k][0][ok]).all(), k
    for k in ("rsi", "natr", "willr", "kdj_k", "kdj_d", "kdj_j", "obv"):
        ok = base[k][1]
        assert T.same_bits(base[k][0][ok], scaled[k][0][ok]).all(), k
    p.close()


def test_single_column_entry_points(pq):
    """pqb_sma / pqb_ema / ... : one reference plugin call on one column through the C ABI."""
    import ctypes as C
    from polars_quant_b200 import _native as N
    L = N.lib()
    eng = pq.get_engine(0)
    d = synth.ohlcv(1, 777, seed=9)
    n = 777

    def col(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return a, N.Col(a.ctypes.data, None, 0, n)

    def outs(k):
        vals = [np.empty(n) for _ in range(k)]
        bits = [np.zeros((n + 7) // 8, np.uint8) for _ in range(k)]
        oc = [N.OutCol(v.ctypes.data, b.ctypes.data) for v, b in zip(vals, bits)]
        return vals, bits, oc

    def check(name, vals, bits, ref):
        ref = ref if isinstance(ref, tuple) and isinstance(ref[0], tuple) else (ref,)
        for v, b, (rv, rk) in zip(vals, bits, ref):
            gok = np.unpackbits(b, bitorder="little")[:n].astype(bool)
            nbad, msg = T.compare(name, v, gok, rv, rk)
            assert nbad == 0, msg

    c_arr, c = col(d["close"][0]); h_arr, h = col(d["high"][0]); l_arr, l = col(d["low"][0]); v_arr, v = col(d["volume"][0])
    vals, bits, oc = outs(1)
    N.check(L.pqb_sma(eng._h, C.byref(c), 10, C.byref(oc[0]))); check("sma", vals, bits, pqo.sma(c_arr, 10))
    N.check(L.pqb_ema(eng._h, C.byref(c), 50, C.byref(oc[0]))); check("ema", vals, bits, pqo.ema(c_arr, 50))
    N.check(L.pqb_tema(eng._h, C.byref(c), 7, C.byref(oc[0]))); check("tema", vals, bits, pqo.tema(c_arr, 7))
    N.check(L.pqb_trima(eng._h, C.byref(c), 9, C.byref(oc[0]))); check("trima", vals, bits, pqo.trima(c_arr, 9))
    N.check(L.pqb_ma(eng._h, C.byref(c), 12, 1, C.byref(oc[0]))); check("ma", vals, bits, pqo.ma(c_arr, 12, 1))
    N.check(L.pqb_rsi(eng._h, C.byref(c), 14, C.byref(oc[0]))); check("rsi", vals, bits, pqo.rsi(c_arr, 14))
    N.check(L.pqb_atr(eng._h, C.byref(h), C.byref(l), C.byref(c), 14, C.byref(oc[0])))
    check("atr", vals, bits, pqo.atr(h_arr, l_arr, c_arr, 14))
    N.check(L.pqb_obv(eng._h, C.byref(c), C.byref(v), C.byref(oc[0]))); check("obv", vals, bits, pqo.obv(c_arr, v_arr))
    N.check(L.pqb_ad(eng._h, C.byref(h), C.byref(l), C.byref(c), C.byref(v), C.byref(oc[0])))
    check("ad", vals, bits, pqo.ad(h_arr, l_arr, c_arr, v_arr))
    N.check(L.pqb_willr(eng._h, C.byref(h), C.byref(l), C.byref(c), 100, C.byref(oc[0])))
    check("willr", vals, bits, pqo.willr(h_arr, l_arr, c_arr, 100))
    N.check(L.pqb_midprice(eng._h, C.byref(h), C.byref(l), 14, C.byref(oc[0])))
    check("midprice", vals, bits, pqo.midprice(h_arr, l_arr, 14))
    vals, bits, oc = outs(3)
    N.check(L.pqb_bbands(eng._h, C.byref(c), 20, C.c_double(2.0), C.c_double(2.0), C.byref(oc[0]), C.byref(oc[1]), C.byref(oc[2])))
    check("bbands", vals, bits, pqo.bbands(c_arr, 20, 2.0, 2.0))
    N.check(L.pqb_macd(eng._h, C.byref(c), 12, 26, 9, C.byref(oc[0]), C.byref(oc[1]), C.byref(oc[2])))
    check("macd", vals, bits, pqo.macd(c_arr, 12, 26, 9))
    N.check(L.pqb_kdj(eng._h, C.byref(h), C.byref(l), C.byref(c), 9, 3, 3, C.byref(oc[0]), C.byref(oc[1]), C.byref(oc[2])))
    check("kdj", vals, bits, pqo.kdj(h_arr, l_arr, c_arr, 9, 3, 3))
    vals, bits, oc = outs(2)
    N.check(L.pqb_stoch(eng._h, C.byref(h), C.byref(l), C.byref(c), 5, 3, 3, C.byref(oc[0]), C.byref(oc[1])))
    check("stoch", vals, bits, pqo.stoch(h_arr, l_arr, c_arr, 5, 3, 0, 3, 0))
    # nulls: momentum.rs functions fail like the reference's cont_slice()?
    bm = np.full((n + 7) // 8, 0xFF, np.uint8); bm[5] = 0xFE
    cn = N.Col(c_arr.ctypes.data, bm.ctypes.data, 0, n)
    vals, bits, oc = outs(1)
    assert L.pqb_rsi(eng._h, C.byref(cn), 14, C.byref(oc[0])) == -5


def test_multi_gpu_driver_shards_by_symbol(pq):
    """pqb_multi_*: the panel split over every visible GPU (two shards on one device when the box has a
    single GPU), one host thread per shard, no collective; every symbol equals the oracle."""
    import ctypes as C
    from polars_quant_b200 import _native as N
    n_dev = max(1, N.lib().pqb_device_count())
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    S, NB = 203, 600
    d = synth.ohlcv(S, NB, seed=88)
    mp = pq.MultiPanel(S, NB, devices)
    sh = mp.shards()
    assert sh[0][1] == 0 and sh[-1][2] == S and all(a[2] == b[1] for a, b in zip(sh, sh[1:]))
    for s in range(S):
        for f in ("close", "high", "low", "volume"):
            mp.set_column(s, f, d[f][s])
    mp.run_host()
    out, ok = _oracle(d)
    for s in range(0, S, 7):
        for k, name in enumerate(pqo.OUTPUT_NAMES):
            v, b = mp.get_output(s, k)
            nbad, msg = T.compare(name, v, b, out[k][s], ok[k][s])
            assert nbad == 0, f"symbol {s}: {msg}"
    mp.close()


@pytest.mark.parametrize("n_symbols", [70, 4736 + 96])
def test_closes_at_the_window_extremes_keep_their_signed_zeros(pq, n_symbols):
    """Real bars close at their high or low now and then: WILLR's and STOCH's numerators are then exactly 0 (and
    -100 * 0 is -0.0): values and signs must equal the oracle bit for bit through the divisions' fast and slow paths --
    plain kernel (70 symbols: three CTAs per SM) and the pipelined small-panel variant (151 blocks)."""
    NB = 300
    d = synth.ohlcv(n_symbols, NB, seed=77)
    rng = np.random.default_rng(3)
    m = rng.random((n_symbols, NB)) < 0.15
    d["high"][m] = d["close"][m]
    m = rng.random((n_symbols, NB)) < 0.15
    d["low"][m] = d["close"][m]
    _run_vs_oracle(pq, d)
    names = ("willr", "kdj_k", "kdj_d", "kdj_j")                 # and as a partial suite (BASE kernel, pipelined roles)
    panel = pq.Panel(n_symbols, NB, outputs_mask=sum(1 << pqo.OUTPUT_NAMES.index(o) for o in names))
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"])
    from polars_quant_b200 import _native as N
    res = panel.compute(N.default_params(indicators=N.IND["willr"] | N.IND["kdj"]))
    out, ok = _oracle(d)
    for name in names:
        k = pqo.OUTPUT_NAMES.index(name)
        nbad, msg = T.compare(name, res[name][0], res[name][1], out[k], ok[k])
        assert nbad == 0, msg
    assert (np.signbit(res["willr"][0]) & (res["willr"][0] == 0.0)).any()      # the -0.0 case occurs
    panel.close()
