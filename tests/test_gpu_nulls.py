"""Null-aware mode of the CUDA path (interior / trailing nulls, fields starting at different rows)
against the oracle's null semantics, function by function, bit-exact (SURVEY.md 8a conventions:
overlap/volatility/volume functions skip nulls, trange/obv use a positional shift, STOCH uses polars'
positional rolling windows, momentum.rs functions and midprice fail -> all-null column)."""
import numpy as np
import pytest

import synth
from oracle import pqo

pytestmark = pytest.mark.gpu

import tolerances as T

F = ("close", "high", "low", "volume")


def _panel_with_nulls(pq, d, ok):
    S, N = d["close"].shape
    panel = pq.Panel(S, N)
    for s in range(S):
        for f in F:
            bits = np.packbits(ok[f][s].astype(np.uint8), bitorder="little")
            panel.set_column(s, f, d[f][s], validity=bits)
    return panel


def _check(name, res, s, ref):
    v, k = ref
    nbad, msg = T.compare(name, res[name][0][s], res[name][1][s], v, k)
    assert nbad == 0, f"symbol {s}: {msg}"


def _null_masks(S, N, seed, frac=0.03):
    rng = np.random.default_rng(seed)
    ok = {f: rng.random((S, N)) >= frac for f in F}
    for f in F:
        ok[f][0] = True                                      # symbol 0: no nulls at all
    lead = rng.integers(0, 60, S)
    for s in range(1, S):
        for f in F:
            ok[f][s, :lead[s]] = False                       # shared leading nulls
    ok["volume"][2, :25] = False                             # a field that starts later than the others
    for f in F:
        ok[f][3, N - 40:] = False                            # delisted: trailing nulls
        ok[f][4] = True
    ok["close"][4, 100] = False                              # exactly one interior null
    for f in F:
        ok[f][5] = True
        ok[f][5, :17] = False                                # leading nulls only, same for all fields
    return ok


_DEFAULT_PERIODS = dict(sma_period=30, ema_period=30, tema_period=30, trima_period=30, bbands_period=20, bbands_nbdevup=2.0, bbands_nbdevdn=2.0,
                        macd_fast=12, macd_slow=26, macd_signal=9, rsi_period=14, atr_period=14, natr_period=14,
                        kdj_fastk=9, kdj_slowk=3, kdj_slowd=3)


def _check_symbol_against_the_oracle(res, s, d, ok, N, P=_DEFAULT_PERIODS):
    c, h, l, v = (d[f] for f in F)
    kc, kh, kl, kv = (ok[f] for f in F)
    _check("sma", res, s, pqo.sma(c[s], P["sma_period"], kc[s]))
    _check("ema", res, s, pqo.ema(c[s], P["ema_period"], kc[s]))
    _check("tema", res, s, pqo.tema(c[s], P["tema_period"], kc[s]))
    _check("trima", res, s, pqo.trima(c[s], P["trima_period"], kc[s]))
    for name, ref in zip(("bb_upper", "bb_middle", "bb_lower"),
                         pqo.bbands(c[s], P["bbands_period"], P["bbands_nbdevup"], P["bbands_nbdevdn"], kc[s])):
        _check(name, res, s, ref)
    _check("trange", res, s, pqo.trange(h[s], l[s], c[s], kh[s], kl[s], kc[s]))
    _check("atr", res, s, pqo.atr(h[s], l[s], c[s], P["atr_period"], kh[s], kl[s], kc[s]))
    _check("natr", res, s, pqo.natr(h[s], l[s], c[s], P["natr_period"], kh[s], kl[s], kc[s]))
    _check("obv", res, s, pqo.obv(c[s], v[s], kc[s], kv[s]))
    _check("ad", res, s, pqo.ad(h[s], l[s], c[s], v[s], kh[s], kl[s], kc[s], kv[s]))
    sk, sd = pqo.stoch(h[s], l[s], c[s], P["kdj_fastk"], P["kdj_slowk"], 0, P["kdj_slowd"], 0, kh[s], kl[s], kc[s])
    _check("kdj_k", res, s, sk)
    _check("kdj_d", res, s, sd)
    jok = sk[1] & sd[1]
    _check("kdj_j", res, s, (np.where(jok, 3.0 * sk[0] - 2.0 * sd[0], np.nan), jok))
    # momentum.rs functions / midprice: the reference fails on interior or trailing nulls
    lead_only = {f: (not ok[f][s].all()) and ok[f][s][np.argmax(ok[f][s]):].all() for f in F}
    clean = {f: ok[f][s].all() or lead_only[f] for f in F}
    a = int(np.argmax(kc[s]))
    if clean["close"]:
        for name, ref in zip(("macd", "macd_signal", "macd_hist"), pqo.macd(c[s, a:], P["macd_fast"], P["macd_slow"], P["macd_signal"])):
            full_v = np.full(N, np.nan); full_k = np.zeros(N, bool)
            full_v[a:], full_k[a:] = ref
            _check(name, res, s, (full_v, full_k))
        rv, rk = pqo.rsi(c[s, a:], P["rsi_period"])
        full_v = np.full(N, np.nan); full_k = np.zeros(N, bool); full_v[a:], full_k[a:] = rv, rk
        _check("rsi", res, s, (full_v, full_k))
    else:
        for name in ("macd", "macd_signal", "macd_hist", "rsi"):
            assert not res[name][1][s].any(), f"{name} symbol {s}: reference fails on nulls -> all null"
    if not (clean["close"] and clean["high"] and clean["low"]):
        assert not res["willr"][1][s].any()
    if not (clean["high"] and clean["low"]):
        assert not res["midprice"][1][s].any()


def test_null_semantics_function_by_function(pq=None):
    import polars_quant_b200 as pq
    S, N = 40, 700
    d = synth.ohlcv(S, N, seed=31)
    ok = _null_masks(S, N, seed=8)
    panel = _panel_with_nulls(pq, d, ok)
    res = panel.compute()
    c, h, l, v = (d[f] for f in F)
    kc, kh, kl, kv = (ok[f] for f in F)
    for s in range(S):
        _check_symbol_against_the_oracle(res, s, d, ok, N)
    # symbols without interior nulls and with one common start are exactly the trimmed series
    for s in (0, 5):
        a = int(np.argmax(kc[s]))
        out, okk, _ = pqo.suite_panel(c[s:s + 1, a:], h[s:s + 1, a:], l[s:s + 1, a:], v[s:s + 1, a:])
        for j, name in enumerate(pqo.OUTPUT_NAMES):
            full_v = np.full(N, np.nan); full_k = np.zeros(N, bool)
            full_v[a:], full_k[a:] = out[j, 0], okk[j, 0]
            _check(name, res, s, (full_v, full_k))
    panel.close()


def test_null_mode_host_pipeline_and_plain_mode_agree():
    """A panel whose only nulls are shared leading nulls gives the same bytes whether it runs in the
    plain mode (starts) or is forced through the null-aware kernel by one extra interior null in a
    different symbol; run_host == compute in null-aware mode."""
    import polars_quant_b200 as pq
    S, N = 70, 400
    d = synth.ohlcv(S, N, seed=13)
    ok = {f: np.ones((S, N), bool) for f in F}
    for f in F:
        ok[f][7, :33] = False
    plain = _panel_with_nulls(pq, d, ok)
    r_plain = {k: (a.copy(), b.copy()) for k, (a, b) in plain.compute().items()}
    ok["close"][50, 200] = False
    nul = _panel_with_nulls(pq, d, ok)
    r_null = {k: (a.copy(), b.copy()) for k, (a, b) in nul.compute().items()}
    for name in r_plain:
        for s in list(range(0, 50)) + list(range(51, S)):
            assert np.array_equal(r_plain[name][1][s], r_null[name][1][s]), (name, s)
            assert T.same_bits(r_plain[name][0][s], r_null[name][0][s]).all(), (name, s)
    nul.run_host(chunk_symbols=32)
    r_host = nul.outputs()
    for name in r_null:
        assert np.array_equal(r_null[name][1], r_host[name][1]), name
        assert T.same_bits(r_null[name][0], r_host[name][0]).all(), name
    plain.close(); nul.close()


def test_single_column_calls_with_nulls():
    """pqb_sma / pqb_ema / pqb_atr / pqb_obv skip nulls like the reference; pqb_rsi / pqb_macd refuse."""
    import ctypes as C
    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as N
    L = N.lib()
    eng = pq.get_engine(0)
    n = 500
    d = synth.ohlcv(1, n, seed=99)
    rng = np.random.default_rng(3)
    ok = rng.random(n) > 0.05
    bits = np.packbits(ok.astype(np.uint8), bitorder="little")
    c = np.ascontiguousarray(d["close"][0]); h = np.ascontiguousarray(d["high"][0]); l = np.ascontiguousarray(d["low"][0])
    v = np.ascontiguousarray(d["volume"][0])
    col = lambda a, b=None: N.Col(a.ctypes.data, None if b is None else b.ctypes.data, 0, n)
    out_v, out_b = np.empty(n), np.zeros((n + 7) // 8, np.uint8)
    oc = N.OutCol(out_v.ctypes.data, out_b.ctypes.data)

    def got():
        return out_v.copy(), np.unpackbits(out_b, bitorder="little")[:n].astype(bool)

    cc = col(c, bits)
    N.check(L.pqb_sma(eng._h, C.byref(cc), 10, C.byref(oc)))
    assert T.compare("sma", *got(), *pqo.sma(c, 10, ok))[0] == 0
    N.check(L.pqb_ema(eng._h, C.byref(cc), 10, C.byref(oc)))
    assert T.compare("ema", *got(), *pqo.ema(c, 10, ok))[0] == 0
    ch, cl = col(h), col(l)
    N.check(L.pqb_atr(eng._h, C.byref(ch), C.byref(cl), C.byref(cc), 14, C.byref(oc)))
    assert T.compare("atr", *got(), *pqo.atr(h, l, c, 14, None, None, ok))[0] == 0
    cv = col(v)
    N.check(L.pqb_obv(eng._h, C.byref(cc), C.byref(cv), C.byref(oc)))
    assert T.compare("obv", *got(), *pqo.obv(c, v, ok, None))[0] == 0
    assert L.pqb_rsi(eng._h, C.byref(cc), 14, C.byref(oc)) == -5
    # and a clean call right after a null-aware one (scratch panel state is reset)
    N.check(L.pqb_sma(eng._h, C.byref(col(c)), 10, C.byref(oc)))
    assert T.compare("sma", *got(), *pqo.sma(c, 10))[0] == 0


def _partial_suite_still_compacts(panel, ref):
    """The full suite sends flagged blocks through its specialised null-aware kernel (per-block dispatch, engine.cu launch_suite);
    symbol compaction remains the path of every other launch: a partial suite on the same device-resident panel must reproduce
    the full run's values for the groups it computes (`ref` = the chunked host pipeline's outputs of the full suite)."""
    from polars_quant_b200 import _native as Nn
    skip = {"willr", "kdj_k", "kdj_d", "kdj_j"}
    ref = {k: (v[0].copy(), v[1].copy()) for k, v in ref.items()}
    panel.upload()
    panel.run(Nn.default_params(indicators=Nn.IND_ALL & ~(Nn.IND["willr"] | Nn.IND["kdj"])))
    panel.download()
    panel.sync()
    got = panel.outputs()
    for name in pqo.OUTPUT_NAMES:
        if name in skip:
            continue
        assert np.array_equal(got[name][1], ref[name][1]), name
        m = ref[name][1]
        assert np.array_equal(got[name][0][m].view(np.uint64), ref[name][0][m].view(np.uint64)), name


def test_symbol_compaction_few_flagged_symbols_in_a_large_panel():
    """A device-resident panel in which FEW symbols need the null-aware kernel (halts, a delisting, a field that starts late)
    and they are spread over many blocks: the engine copies them into blocks of their own, runs the null-aware kernel on those
    beside the plain kernel on every original block, and writes their lanes back (engine.cu "symbol compaction").  Every output
    of every flagged symbol and of its block neighbours against the oracle; the whole panel against the chunked host pipeline,
    which never compacts (per-block dispatch), bit for bit."""
    import polars_quant_b200 as pq
    S, N = 1500, 520
    d = synth.ohlcv(S, N, seed=77)
    ok = {f: np.ones((S, N), bool) for f in F}
    rng = np.random.default_rng(3)
    lead = rng.integers(0, 40, S)
    for s in range(S):
        if s % 7 == 0:
            for f in F:
                ok[f][s, :lead[s]] = False                    # later listings: leading nulls shared by the fields (plain path)
    flagged = [5, 70, 131, 200, 333, 334, 470, 512, 777, 900, 1023, 1250, 1499]
    for i, s in enumerate(flagged):
        kind = i % 5
        if kind == 0: ok["close"][s, 200:203] = False         # a halt in close
        elif kind == 1:
            for f in F: ok[f][s, N - 60:] = False             # delisted
        elif kind == 2: ok["volume"][s, :33] = False          # volume starts later than the other fields
        elif kind == 3: ok["high"][s, 301] = False            # one missing high
        else:
            for f in F: ok[f][s, 100:140] = False             # a long halt in every field
    panel = _panel_with_nulls(pq, d, ok)
    res = panel.compute()
    for s in sorted(set(flagged + [4, 6, 69, 71, 332, 335, 1498, 0, 1, 7, 14])):
        _check_symbol_against_the_oracle(res, s, d, ok, N)
    got = {k: (v[0].copy(), v[1].copy()) for k, v in res.items()}
    panel.run_host(chunk_symbols=32)                          # one launch per 32-symbol chunk: per-block dispatch, no compaction
    ref = panel.outputs()
    for name in pqo.OUTPUT_NAMES:
        assert np.array_equal(got[name][1], ref[name][1]), name
        m = ref[name][1]
        assert np.array_equal(got[name][0][m].view(np.uint64), ref[name][0][m].view(np.uint64)), name
    _partial_suite_still_compacts(panel, ref)
    panel.close()


def test_symbol_compaction_direct_mode_many_flagged_symbols():
    """More than 1,024 flagged symbols (a third of the panel: delistings): the null-aware kernel runs after the plain one and stores
    straight into the flagged symbols' own lanes (no write-back pass).  Sampled symbols against the oracle; the whole panel against
    the chunked host pipeline (per-block dispatch), bit for bit."""
    import polars_quant_b200 as pq
    S, N = 3400, 260
    d = synth.ohlcv(S, N, seed=91)
    ok = {f: np.ones((S, N), bool) for f in F}
    rng = np.random.default_rng(12)
    flagged = sorted(rng.choice(S, size=1100, replace=False).tolist())
    for i, s in enumerate(flagged):
        kind = i % 4
        if kind == 0:
            for f in F: ok[f][s, N - 20 - (i % 50):] = False      # delisted at different dates
        elif kind == 1: ok["close"][s, 100 + (i % 40)] = False    # one missing close
        elif kind == 2: ok["low"][s, 50:53] = False
        else: ok["volume"][s, :10 + (i % 7)] = False              # volume starts later
    for s in range(0, S, 11):
        if s not in flagged:
            for f in F: ok[f][s, :s % 37] = False                 # leading nulls shared by the fields (plain path)
    panel = _panel_with_nulls(pq, d, ok)
    res = panel.compute()
    sample = flagged[:6] + flagged[-6:] + [0, 1, 11, 22, S - 1] + [s + 1 for s in flagged[100:104] if s + 1 < S]
    for s in sorted(set(sample)):
        _check_symbol_against_the_oracle(res, s, d, ok, N)
    got = {k: (v[0].copy(), v[1].copy()) for k, v in res.items()}
    panel.run_host(chunk_symbols=32)
    ref = panel.outputs()
    for name in pqo.OUTPUT_NAMES:
        assert np.array_equal(got[name][1], ref[name][1]), name
        m = ref[name][1]
        assert np.array_equal(got[name][0][m].view(np.uint64), ref[name][0][m].view(np.uint64)), name
    _partial_suite_still_compacts(panel, ref)
    panel.close()


def test_a_reused_panel_forgets_the_nulls_of_overwritten_fields():
    """A panel that held columns with interior nulls and is then refilled with dense matrices (Panel.set_fields writes the
    pinned planes directly) must not stay in null-aware mode on stale validity: every output equals the plain oracle and the
    optional groups, which the null-aware kernel refuses, run."""
    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as NV
    S, N = 40, 300
    d = synth.ohlcv(S, N, seed=5)
    ok = _null_masks(S, N, seed=2)
    panel = _panel_with_nulls(pq, d, ok)
    panel.compute()                                            # null-aware run
    d2 = synth.ohlcv(S, N, seed=6)
    panel.set_fields(close=d2["close"], high=d2["high"], low=d2["low"], volume=d2["volume"])
    res = panel.compute()
    out, okk, _ = pqo.suite_panel(d2["close"], d2["high"], d2["low"], d2["volume"])
    fails = T.compare_all(res, out, okk, pqo.OUTPUT_NAMES)
    assert not fails, "\n".join(fails[:5])
    panel.close()
    p2 = pq.Panel(S, N, outputs_mask=(1 << NV.N_SUITE_OUTPUTS) - 1 | 1 << NV.OUTPUT_NAMES.index("mom"))
    for s in range(S):
        bits = np.packbits(ok["close"][s].astype(np.uint8), bitorder="little")
        p2.set_column(s, "close", d["close"][s], validity=bits)
    p2.set_fields(close=d2["close"], high=d2["high"], low=d2["low"], volume=d2["volume"])
    res = p2.compute(NV.default_params(indicators=NV.IND_ALL | NV.IND_EXTRA["mom"]))     # refused on a panel with interior nulls
    mv, mk = pqo.mom(d2["close"][3], 10)
    nbad, msg = T.compare("mom", res["mom"][0][3], res["mom"][1][3], mv, mk)
    assert nbad == 0, msg
    p2.close()


@pytest.mark.parametrize("seed", range(3))
def test_null_mode_fast_stages_between_clustered_halts(seed):
    """Halts come in clusters (a market-wide suspension, a data outage) with hundreds of clean bars between them: stages in which
    every lane of a block is valid, after `steady_lead` such bars in a row, run the plain steady step inside the null-aware kernel
    (suite_kernel.cuh run_role "fast stage") and the per-function counters move by whole stages.  Every transition -- slow -> fast
    after each cluster, fast -> slow at the next one, the delisting tail, a symbol without any null in a flagged block, fields
    that halt alone -- against the oracle, every symbol, default and random periods."""
    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as Nn
    rng = np.random.default_rng(9100 + seed)
    S, N = 72, 1500
    d = synth.ohlcv(S, N, seed=810 + seed)
    ok = {f: np.ones((S, N), bool) for f in F}
    clusters = [(290, 330), (771, 779), (1203, 1204)]
    for s in range(S):
        if s % 9 == 8:
            continue                                          # no null at all, in a block whose other lanes have some
        for lo, hi in clusters:
            if rng.random() < 0.25:
                continue
            a = int(rng.integers(lo, hi)); n = int(rng.integers(1, 6))
            fields = [F[i] for i in range(4) if rng.random() < 0.4] or [F[int(rng.integers(0, 4))]]
            for f in fields:
                ok[f][s, a:a + n] = False
        if s % 5 == 0:
            for f in F: ok[f][s, :int(rng.integers(1, 50))] = False      # listed later
        if s % 13 == 3:
            for f in F: ok[f][s, N - int(rng.integers(5, 120)):] = False  # delisted
    P = dict(_DEFAULT_PERIODS)
    if seed:
        r = lambda lo, hi: int(rng.integers(lo, hi + 1))
        fast = r(1, 15)
        P = dict(sma_period=r(1, 60), ema_period=r(1, 60), tema_period=r(1, 25), trima_period=r(1, 60), bbands_period=r(1, 50),
                 bbands_nbdevup=1.5, bbands_nbdevdn=2.5, macd_fast=fast, macd_slow=fast + r(1, 20), macd_signal=r(1, 12), rsi_period=r(1, 30),
                 atr_period=r(1, 30), natr_period=r(1, 30), kdj_fastk=r(1, 40), kdj_slowk=r(1, 8), kdj_slowd=r(1, 8))
    panel = _panel_with_nulls(pq, d, ok)
    res = panel.compute(Nn.default_params(**P))
    for s in range(S):
        _check_symbol_against_the_oracle(res, s, d, ok, N, P)
    got = {k: (v[0].copy(), v[1].copy()) for k, v in res.items()}
    panel.run_host(Nn.default_params(**P), chunk_symbols=32)   # per-block dispatch instead of symbol compaction
    ref = panel.outputs()
    for name in pqo.OUTPUT_NAMES:
        assert np.array_equal(got[name][1], ref[name][1]), name
        m = ref[name][1]
        assert np.array_equal(got[name][0][m].view(np.uint64), ref[name][0][m].view(np.uint64)), name
    panel.close()


@pytest.mark.parametrize("seed", range(6))
def test_null_mode_random_periods_and_null_patterns(seed):
    """Random periods (1 included) x random null patterns: the null-aware walk keeps its per-function rules whatever the window
    lengths are -- windows shorter than a gap, longer than the listed history, period-1 quirks."""
    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as Nn
    rng = np.random.default_rng(4200 + seed)
    r = lambda lo, hi: int(rng.integers(lo, hi + 1))
    fast = r(1, 15)
    P = dict(sma_period=r(1, 60), ema_period=r(1, 60), tema_period=r(1, 25), trima_period=r(1, 60), bbands_period=r(1, 50),
             bbands_nbdevup=1.5, bbands_nbdevdn=2.5, macd_fast=fast, macd_slow=fast + r(1, 20), macd_signal=r(1, 12), rsi_period=r(1, 30),
             atr_period=r(1, 30), natr_period=r(1, 30), kdj_fastk=r(1, 40), kdj_slowk=r(1, 8), kdj_slowd=r(1, 8))
    if seed % 2:
        P["natr_period"] = P["atr_period"]
        P["ema_period"] = P["tema_period"]
    S, N = 40, 500
    d = synth.ohlcv(S, N, seed=700 + seed)
    ok = _null_masks(S, N, seed=50 + seed, frac=(0.002, 0.03, 0.15)[seed % 3])
    panel = _panel_with_nulls(pq, d, ok)
    res = panel.compute(Nn.default_params(**P))
    for s in range(S):
        _check_symbol_against_the_oracle(res, s, d, ok, N, P)
    panel.close()
