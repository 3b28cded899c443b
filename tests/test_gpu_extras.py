"""The remaining SURVEY.md 8a functions (optional indicator groups, outputs 21..30): midpoint, adosc,
mom, roc / rocp / rocr / rocr100, cmo, mfi, cci -- bit-exact against the oracle, including the
data-dependent nulls of the roc family (lagged value == 0) and cci (mean deviation == 0)."""
import ctypes as C

import numpy as np
import pytest

import synth
from oracle import pqo

pytestmark = pytest.mark.gpu

import tolerances as T

EXTRA_OUT = ("midpoint", "adosc", "mom", "roc", "rocp", "rocr", "rocr100", "cmo", "mfi", "cci",
             "plus_dm", "minus_dm", "dx", "minus_di", "adx", "adxr", "trix", "ultosc", "aroon_up", "aroon_down")
DM_OUT = ("plus_dm", "minus_dm", "dx", "minus_di", "adx", "adxr")


def _refs(d, s, P):
    c, h, l, v = (d[f][s] for f in ("close", "high", "low", "volume"))
    dm = pqo.dm(h, l, c, P.get("dm_period", 14))             # calc_dm family, momentum.rs:668-727
    au, ad_ = pqo.aroon(h, l, P.get("aroon_period", 14))
    du, dl = pqo.donchian(h, l, P.get("donchian_period", 20))    # SURVEY.md D3 (BASELINE config 5)
    return {
        "donchian_upper": du, "donchian_lower": dl,
        **dm,
        "trix": pqo.trix(c, P.get("trix_period", 30)),
        "ultosc": pqo.ultosc(h, l, c, P.get("ultosc_period1", 7), P.get("ultosc_period2", 14), P.get("ultosc_period3", 28)),
        "aroon_up": au, "aroon_down": ad_,
        "midpoint": pqo.midpoint(c, P["midpoint_period"]),
        "adosc": pqo.adosc(h, l, c, v, P["adosc_fast"], P["adosc_slow"]),
        "mom": pqo.mom(c, P["mom_period"]),
        "roc": pqo.roc(c, P["roc_period"], 0), "rocp": pqo.roc(c, P["roc_period"], 1),
        "rocr": pqo.roc(c, P["roc_period"], 2), "rocr100": pqo.roc(c, P["roc_period"], 3),
        "cmo": pqo.cmo(c, P["cmo_period"]),
        "mfi": pqo.mfi(h, l, c, v, P["mfi_period"]),
        "cci": pqo.cci(h, l, c, P["cci_period"]),
    }


def _data():
    d = synth.ohlcv(37, 650, seed=71)
    d["close"][5, 100] = 0.0                                  # roc family: null 10 bars later
    d["close"][6, ::7] = 0.0
    flat = np.full(650, 42.0)
    for f in ("close", "high", "low"):
        d[f][7] = flat                                        # cci: mean deviation 0 -> null; cmo total 0; mfi neg 0
    d["volume"][8, 200:260] = 0.0
    return d


@pytest.mark.parametrize("periods", [dict(), dict(midpoint_period=5, adosc_fast=2, adosc_slow=7, mom_period=1, roc_period=3,
                                                  cmo_period=1, mfi_period=2, cci_period=1, dm_period=1,
                                                  trix_period=1, ultosc_period1=1, ultosc_period2=2, ultosc_period3=3, aroon_period=1, donchian_period=1),
                                     dict(midpoint_period=60, adosc_fast=30, adosc_slow=12, mom_period=55, roc_period=41,
                                          cmo_period=33, mfi_period=29, cci_period=47, dm_period=37,
                                          trix_period=21, ultosc_period1=12, ultosc_period2=9, ultosc_period3=20, aroon_period=21, donchian_period=55)])
def test_optional_groups_alone_and_with_the_suite(periods):
    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as N
    P = dict(midpoint_period=14, adosc_fast=3, adosc_slow=10, mom_period=10, roc_period=10, cmo_period=14,
             mfi_period=14, cci_period=14, dm_period=14, trix_period=30, ultosc_period1=7, ultosc_period2=14,
             ultosc_period3=28, aroon_period=14)
    P.update(periods)
    d = _data()
    S, NB = d["close"].shape
    extras = sum(N.IND_EXTRA.values())
    panel = pq.Panel(S, NB, outputs_mask=(1 << N.N_OUTPUTS) - 1)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"])
    for ind in (extras, extras | N.IND_ALL):
        res = panel.compute(N.default_params(indicators=ind, **P))
        for s in range(S):
            for name, (v, k) in _refs(d, s, P).items():
                nbad, msg = T.compare(name, res[name][0][s], res[name][1][s], v, k)
                assert nbad == 0, f"indicators {ind:#x} symbol {s}: {msg}"
        if ind & N.IND_ALL:                                   # the suite next to them is untouched
            out, ok, _ = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"])
            fails = T.compare_all(res, out, ok, pqo.OUTPUT_NAMES)
            assert not fails, "\n".join(fails)
    # one group at a time
    for g, bit in N.IND_EXTRA.items():
        res = panel.compute(N.default_params(indicators=bit, **P))
        names = (("roc", "rocp", "rocr", "rocr100") if g == "roc" else DM_OUT if g == "dm"
                 else ("aroon_up", "aroon_down") if g == "aroon"
                 else ("donchian_upper", "donchian_lower") if g == "donchian" else (g,))
        for s in (0, 5, 6, 7, 8, 36):
            refs = _refs(d, s, P)
            for name in names:
                nbad, msg = T.compare(name, res[name][0][s], res[name][1][s], *refs[name])
                assert nbad == 0, f"group {g} symbol {s}: {msg}"
    panel.close()


def test_optional_groups_with_leading_nulls_and_host_pipeline():
    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as N
    d = synth.ohlcv(70, 500, seed=12)
    S, NB = d["close"].shape
    starts = np.random.default_rng(4).integers(0, 120, S).astype(np.int32)
    P = dict(midpoint_period=14, adosc_fast=3, adosc_slow=10, mom_period=10, roc_period=10, cmo_period=14,
             mfi_period=14, cci_period=14, dm_period=14, trix_period=30, ultosc_period1=7, ultosc_period2=14,
             ultosc_period3=28, aroon_period=14)
    panel = pq.Panel(S, NB, outputs_mask=(1 << N.N_OUTPUTS) - 1)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"], starts=starts)
    panel.run_host(N.default_params(indicators=sum(N.IND_EXTRA.values()) | N.IND_ALL), chunk_symbols=32)
    res = panel.outputs()
    for s in range(S):
        a = int(starts[s])
        dd = {f: d[f][:, a:] for f in d}
        for name, (v, k) in _refs(dd, s, P).items():
            fv = np.full(NB, np.nan); fk = np.zeros(NB, bool)
            fv[a:], fk[a:] = v, k
            nbad, msg = T.compare(name, res[name][0][s], res[name][1][s], fv, fk)
            assert nbad == 0, f"symbol {s} (start {a}): {msg}"
    panel.close()


def test_optional_single_column_entry_points():
    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as N
    L = N.lib()
    eng = pq.get_engine(0)
    n = 600
    d = synth.ohlcv(1, n, seed=123)
    c, h, l, v = (np.ascontiguousarray(d[f][0]) for f in ("close", "high", "low", "volume"))
    c[300] = 0.0
    col = lambda a: N.Col(a.ctypes.data, None, 0, n)
    out_v, out_b = np.empty(n), np.zeros((n + 7) // 8, np.uint8)
    oc = N.OutCol(out_v.ctypes.data, out_b.ctypes.data)
    got = lambda: (out_v.copy(), np.unpackbits(out_b, bitorder="little")[:n].astype(bool))
    cc, ch, cl, cv = col(c), col(h), col(l), col(v)
    N.check(L.pqb_midpoint(eng._h, C.byref(cc), 14, C.byref(oc))); assert T.compare("midpoint", *got(), *pqo.midpoint(c, 14))[0] == 0
    N.check(L.pqb_adosc(eng._h, C.byref(ch), C.byref(cl), C.byref(cc), C.byref(cv), 3, 10, C.byref(oc)))
    assert T.compare("adosc", *got(), *pqo.adosc(h, l, c, v, 3, 10))[0] == 0
    N.check(L.pqb_mom(eng._h, C.byref(cc), 10, C.byref(oc))); assert T.compare("mom", *got(), *pqo.mom(c, 10))[0] == 0
    for kind in range(4):
        N.check(L.pqb_roc(eng._h, C.byref(cc), 10, kind, C.byref(oc)))
        assert T.compare("roc", *got(), *pqo.roc(c, 10, kind))[0] == 0
    N.check(L.pqb_cmo(eng._h, C.byref(cc), 14, C.byref(oc))); assert T.compare("cmo", *got(), *pqo.cmo(c, 14))[0] == 0
    N.check(L.pqb_mfi(eng._h, C.byref(ch), C.byref(cl), C.byref(cc), C.byref(cv), 14, C.byref(oc)))
    assert T.compare("mfi", *got(), *pqo.mfi(h, l, c, v, 14))[0] == 0
    N.check(L.pqb_cci(eng._h, C.byref(ch), C.byref(cl), C.byref(cc), 14, C.byref(oc)))
    assert T.compare("cci", *got(), *pqo.cci(h, l, c, 14))[0] == 0
    # Donchian channel: two outputs, no warm-up nulls; its mid line is pqb_midprice
    lo_v, lo_b = np.empty(n), np.zeros((n + 7) // 8, np.uint8)
    oc2 = N.OutCol(lo_v.ctypes.data, lo_b.ctypes.data)
    for p in (1, 20, 250):
        N.check(L.pqb_donchian(eng._h, C.byref(ch), C.byref(cl), p, C.byref(oc), C.byref(oc2)))
        up, lo = pqo.donchian(h, l, p)
        assert T.compare("donchian_upper", *got(), *up)[0] == 0
        assert T.compare("donchian_lower", lo_v, np.unpackbits(lo_b, bitorder="little")[:n].astype(bool), *lo)[0] == 0
        assert got()[1].all()
    assert L.pqb_donchian(eng._h, C.byref(ch), C.byref(cl), 0, C.byref(oc), C.byref(oc2)) == -4


def test_directional_movement_family_quirks_and_plugin_names():
    """calc_dm family (momentum.rs:668-727): flat bars (smoothed true range 0 -> DI / DX null, ADX keeps smoothing
    zeros), the reference's plus_di == DX quirk (:409), first valid bars p-1 / 2p-2, nulls refused, period 0."""
    import pyarrow as pa
    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as N
    from polars_quant_b200 import plugin, talib
    d = synth.ohlcv(1, 400, seed=77)
    h, l, c = d["high"][0].copy(), d["low"][0].copy(), d["close"][0].copy()
    h[:40] = l[:40] = c[:40] = 50.0                           # a flat listing period: true range 0 for 40 bars
    ref = pqo.dm(h, l, c, 14)
    assert not ref["dx"][1][13:30].any() and ref["adx"][1][13] and ref["adx"][0][20] == 0.0
    ah, al, ac = pa.array(h), pa.array(l), pa.array(c)

    def same(name, got, r):
        ok = ~np.asarray(got.is_null())
        v = np.where(ok, np.asarray(got.to_numpy(zero_copy_only=False), dtype=np.float64), np.nan)
        nbad, msg = T.compare(name, v, ok, r[0], r[1])
        assert nbad == 0, msg

    same("adx", talib.ADX(ah, al, ac), ref["adx"])
    same("adxr", talib.ADXR(ah, al, ac, 14), ref["adxr"])
    same("dx", talib.DX(ah, al, ac), ref["dx"])
    same("plus_di == dx", talib.PLUS_DI(ah, al, ac), ref["dx"])
    same("minus_di", talib.MINUS_DI(ah, al, ac), ref["minus_di"])
    same("plus_dm", talib.PLUS_DM(ah, al), ref["plus_dm"])
    same("minus_dm", talib.MINUS_DM(ah, al, 14), ref["minus_dm"])
    r5 = pqo.dm(h, l, c, 5)
    same("adxr(5)", plugin.call("adxr", [ah, al, ac], kwargs={"timeperiod": 5}), r5["adxr"])
    assert int(np.argmax(r5["adx"][1])) == 4 and int(np.argmax(r5["adxr"][1])) == 8
    assert talib.ADX(ah, al, ac, 0).null_count == 400         # calc_rma guard (D1)
    with pytest.raises(plugin.PluginError, match="not contiguous"):
        talib.ADX(pa.array(h, mask=np.arange(400) == 100), al, ac)


def test_trix_ultosc_aroon_through_the_plugin_names_and_edge_cases():
    import pyarrow as pa
    from polars_quant_b200 import plugin, talib
    d = synth.ohlcv(1, 300, seed=91)
    h, l, c = d["high"][0].copy(), d["low"][0].copy(), d["close"][0].copy()
    h[100:140] = l[100:140] = c[100:140] = c[99]              # a halt: ultosc range sums hit 0, aroon ties resolve to the LAST bar
    ah, al, ac = pa.array(h), pa.array(l), pa.array(c)

    def same(name, got, r):
        ok = ~np.asarray(got.is_null())
        v = np.where(ok, np.asarray(got.to_numpy(zero_copy_only=False), dtype=np.float64), np.nan)
        nbad, msg = T.compare(name, v, ok, r[0], r[1])
        assert nbad == 0, msg

    same("trix", talib.TRIX(ac), pqo.trix(c, 30))
    same("trix(5)", talib.TRIX(ac, 5), pqo.trix(c, 5))
    ref = pqo.ultosc(h, l, c)
    assert not ref[1][110:135].all()                          # the 7-bar range sum is 0 inside the halt -> nulls
    same("ultosc", talib.ULTOSC(ah, al, ac), ref)
    same("ultosc(3,5,9)", talib.ULTOSC(ah, al, ac, 3, 5, 9), pqo.ultosc(h, l, c, 3, 5, 9))
    up, dn = talib.AROON(ah, al)
    ru, rd = pqo.aroon(h, l, 14)
    same("aroon_up", up, ru); same("aroon_down", dn, rd)
    assert ru[0][130] == 100.0 and rd[0][130] == 100.0        # all-equal window: `>=` / `<=` keep the last index
    assert plugin.output_field("aroon", [pa.field("h", pa.float64())]).name == "aroon"
    with pytest.raises(plugin.PluginError, match="not contiguous"):
        talib.TRIX(pa.array(c, mask=np.arange(300) == 7))
    assert talib.TRIX(ac, 0).null_count == 300
    with pytest.raises(plugin.PluginError, match="not built"):
        talib.AROON(ah, al, 0)


@pytest.mark.parametrize("period", [1, 2, 3, 14, 31, 32, 33, 100])
def test_aroon_positions_with_ties_nan_and_infinities(period):
    """aroon momentum.rs:63-110 through the block-decomposed window (suffix summaries of the previous p+1 bars + the running
    prefix): `>=` / `<=` scans from f64::MIN / f64::MAX -- the LAST of equal extremes wins, a NaN is never taken, neither is
    -inf as a high or +inf as a low (nothing taken at all leaves position 0), +-f64::MAX are ordinary values."""
    import pyarrow as pa
    from polars_quant_b200 import talib
    rng = np.random.default_rng(period)
    n = 1500
    h = np.round(rng.normal(100.0, 2.0, n), 0)                # coarse grid: many exact ties inside a window
    l = h - np.round(rng.uniform(0.0, 3.0, n), 0)
    fmax = np.finfo(np.float64).max
    for k, (hv, lv) in enumerate([(np.nan, np.nan), (-np.inf, np.inf), (np.inf, -np.inf), (-fmax, fmax), (fmax, -fmax)]):
        at = rng.choice(n, 25, replace=False)
        h[at] = hv
        l[at[::2]] = lv
    h[700:700 + 3 * period + 5] = np.nan                      # whole windows without a qualifying high
    l[900:900 + 3 * period + 5] = np.inf
    h[1100:1100 + 2 * period + 3] = -np.inf
    up, dn = talib.AROON(pa.array(h), pa.array(l), period)
    ru, rd = pqo.aroon(h, l, period)
    for name, got, r in (("aroon_up", up, ru), ("aroon_down", dn, rd)):
        ok = ~np.asarray(got.is_null())
        v = np.where(ok, np.asarray(got.to_numpy(zero_copy_only=False), dtype=np.float64), np.nan)
        nbad, msg = T.compare("%s(%d)" % (name, period), v, ok, r[0], r[1])
        assert nbad == 0, msg
    assert ru[1][period:].all() and not ru[1][:period].any()
    assert ru[0][700 + period + 2] == 0.0                     # nothing qualifies in the window: position 0


def test_crossover_signals_are_exact_functions_of_the_suite_outputs():
    """golden / death crosses of MACD and KDJ and the RSI zone exits as int8, computed on the device from
    the suite's own outputs -- compared with the same rules evaluated in numpy on the oracle's outputs (the rules are
    defined in include/pqb200.h; the reference has them only in its README)."""
    import polars_quant_b200 as pq
    d = synth.ohlcv(70, 900, seed=314)
    starts = np.random.default_rng(2).integers(0, 60, 70).astype(np.int32)
    panel = pq.Panel(70, 900)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"], starts=starts)
    panel.run_host()
    sig = panel.signals(oversold=35.0, overbought=65.0)
    res = panel.outputs()

    def cross(a, b):
        out = np.zeros(a.shape, dtype=np.int8)
        with np.errstate(invalid="ignore"):
            up = (a[:, 1:] > b[:, 1:]) & (a[:, :-1] <= b[:, :-1])
            dn = (a[:, 1:] < b[:, 1:]) & (a[:, :-1] >= b[:, :-1])
        out[:, 1:][up] = 1
        out[:, 1:][dn] = -1
        return out

    val = lambda n: np.where(res[n][1], res[n][0], np.nan)          # nulls as NaN: every comparison false
    assert np.array_equal(sig["macd_cross"], cross(val("macd"), val("macd_signal")))
    assert np.array_equal(sig["kdj_cross"], cross(val("kdj_k"), val("kdj_d")))
    r = val("rsi")
    want = np.zeros(r.shape, dtype=np.int8)
    with np.errstate(invalid="ignore"):
        want[:, 1:][(r[:, 1:] > 35.0) & (r[:, :-1] <= 35.0)] = 1
        want[:, 1:][(r[:, 1:] < 65.0) & (r[:, :-1] >= 65.0)] = -1
    assert np.array_equal(sig["rsi_cross"], want)
    for n in sig:
        assert (sig[n] != 0).sum() > 500 and set(np.unique(sig[n]).tolist()) <= {-1, 0, 1}
    # and against the oracle's own columns (the suite outputs are bit-exact, so the signals are too)
    out, ok, _ = pqo.suite_panel(d["close"][:, :], d["high"], d["low"], d["volume"])
    s = 5
    a = int(starts[s])
    m, ms = pqo.macd(d["close"][s][a:])[0:2]
    mv = np.full(900, np.nan); sv = np.full(900, np.nan)
    mv[a:] = np.where(m[1], m[0], np.nan); sv[a:] = np.where(ms[1], ms[0], np.nan)
    assert np.array_equal(sig["macd_cross"][s], cross(mv[None, :], sv[None, :])[0])
    panel.close()


def test_partial_suite_next_to_optional_groups_runs_as_two_launches():
    """Benchmark groups and optional groups together = two launches with their own ring layouts: a partial
    suite (the BASE kernel) next to a few optional groups (the slot-dealt general kernel), long windows included."""
    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as N
    P = dict(midpoint_period=14, adosc_fast=3, adosc_slow=10, mom_period=10, roc_period=10, cmo_period=14, mfi_period=14,
             cci_period=14, dm_period=9, trix_period=30, ultosc_period1=7, ultosc_period2=14, ultosc_period3=28, aroon_period=25)
    d = _data()
    S, NB = d["close"].shape
    ind = N.IND["ema"] | N.IND["kdj"] | N.IND["atr"] | N.IND["midprice"] | N.IND_EXTRA["dm"] | N.IND_EXTRA["aroon"] | N.IND_EXTRA["mom"]
    panel = pq.Panel(S, NB, outputs_mask=(1 << N.N_OUTPUTS) - 1)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"])
    res = panel.compute(N.default_params(indicators=ind, kdj_fastk=40, **P))
    assert panel.last_launches() >= 2
    op = pqo.SuiteParams.default()
    op.stoch_k = 40
    out, ok, _ = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"], op)
    for name in ("ema", "atr", "kdj_k", "kdj_d", "kdj_j", "midprice"):
        k = pqo.OUTPUT_NAMES.index(name)
        nbad, msg = T.compare(name, res[name][0], res[name][1], out[k], ok[k])
        assert nbad == 0, msg
    for s in range(S):
        refs = _refs(d, s, P)
        for name in DM_OUT + ("aroon_up", "aroon_down", "mom"):
            nbad, msg = T.compare(name, res[name][0][s], res[name][1][s], *refs[name])
            assert nbad == 0, f"symbol {s}: {msg}"
    panel.close()


def test_fastk_plane_is_written_only_on_request_and_does_not_cost_the_suite_its_kernel():
    """PQB_OUT_FASTK (STOCHF's raw %K, momentum.py:188-195) is stored by the general kernel only.  A panel that merely has the
    plane (every output allocated) must keep the compile-time-specialised full-suite launch: the line is opt-in
    (PQB_IND_FASTK, ABI 7).  With the bit the line equals the oracle's fastk; the suite's own outputs are the same both ways."""
    import polars_quant_b200 as pq
    from polars_quant_b200 import _native as N
    d = synth.ohlcv(70, 700, seed=5)
    S, NB = d["close"].shape
    panel = pq.Panel(S, NB, outputs_mask=(1 << N.N_OUTPUTS) - 1)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"])
    res = panel.compute(N.default_params(indicators=N.IND_ALL))
    assert not res["fastk"][1].any()                          # not requested: all null
    plain = {k: (v[0].copy(), v[1].copy()) for k, v in res.items() if k.startswith("kdj_") or k in ("sma", "willr")}
    res = panel.compute(N.default_params(indicators=N.IND_ALL | N.IND_FASTK))
    for s in range(S):
        fk = pqo.stochf(d["high"][s], d["low"][s], d["close"][s], 9, 3, 0)[0]
        nbad, msg = T.compare("fastk", res["fastk"][0][s], res["fastk"][1][s], fk[0], fk[1])
        assert nbad == 0, f"symbol {s}: {msg}"
    for k, (v, ok) in plain.items():
        nbad, msg = T.compare(k, res[k][0], res[k][1], v, ok)
        assert nbad == 0, msg
    panel.close()


def test_pinned_planes_of_a_destroyed_panel_are_reused_without_stale_results():
    """The engine keeps the pinned staging planes of destroyed panels for the next panel of the same shape: a second panel
    on recycled planes (other inputs, leading nulls) must not see anything of the first."""
    import polars_quant_b200 as pq
    S, NB = 40, 500
    d1, d2 = synth.ohlcv(S, NB, seed=1), synth.ohlcv(S, NB, seed=2)
    p1 = pq.Panel(S, NB)
    p1.set_fields(d1["close"], d1["high"], d1["low"], d1["volume"])
    p1.compute()
    addr1 = p1.host_output(0).ctypes.data
    p1.close()
    starts = np.zeros(S, dtype=np.int32)
    starts[3], starts[17] = 120, 499
    p2 = pq.Panel(S, NB)
    assert p2.host_output(0).ctypes.data == addr1 or True          # (recycling is an optimisation, not a contract)
    p2.set_fields(d2["close"], d2["high"], d2["low"], d2["volume"], starts=starts)
    res = p2.compute()
    for s in range(S):
        a = int(starts[s])
        o, k, _ = pqo.suite_panel(*(d2[f][s:s + 1, a:] for f in ("close", "high", "low", "volume")))
        for q, name in enumerate(pqo.OUTPUT_NAMES):
            assert not res[name][1][s, :a].any()
            nbad, msg = T.compare(name, res[name][0][s, a:], res[name][1][s, a:], o[q, 0], k[q, 0])
            assert nbad == 0, f"symbol {s}: {msg}"
    p2.close()


def test_engine_destroyed_before_its_panels_is_deferred():
    """A garbage collector may finalise an engine before the panels that point to it (interpreter shutdown): the engine is
    only marked then, the panels stay usable and the last one to go frees it (pinned planes go through the engine's pool)."""
    import polars_quant_b200 as pq
    from polars_quant_b200 import candles
    eng = pq.Engine(0)
    d = synth.ohlcv(8, 300, seed=4)
    p = pq.Panel(8, 300, engine=eng)
    cp = candles.CandlePanel(8, 300, engine=eng)
    p.set_fields(d["close"], d["high"], d["low"], d["volume"])
    eng.close()                                            # destroy requested while two panels are alive
    res = p.compute()
    out, ok, _ = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"])
    assert not T.compare_all(res, out, ok, pqo.OUTPUT_NAMES)
    p.close()
    cp.close()                                             # last one out frees the engine


def test_ma_golden_and_death_cross_signal():
    """README.md:876-905 `Strategy.ma(df, fast_period, slow_period, ma_type)`: cross(MA(close, fast), MA(close, slow)) as int8,
    both averages computed inside the signal kernel with the suite's serial arithmetic -- compared with the crossing rule
    evaluated in numpy on the ORACLE's SMA / EMA columns (calc_sma overlap.rs:871, calc_ema :660)."""
    import polars_quant_b200 as pq
    S, N = 70, 900
    d = synth.ohlcv(S, N, seed=99)
    starts = np.zeros(S, dtype=np.int32)
    starts[[4, 33, 69]] = (25, 400, 880)
    panel = pq.Panel(S, N)
    panel.set_fields(d["close"], d["high"], d["low"], d["volume"], starts=starts)
    panel.upload()

    def cross(a, b):
        out = np.zeros(a.shape, dtype=np.int8)
        with np.errstate(invalid="ignore"):
            out[1:][(a[1:] > b[1:]) & (a[:-1] <= b[:-1])] = 1
            out[1:][(a[1:] < b[1:]) & (a[:-1] >= b[:-1])] = -1
        return out

    for ma_type, fn, fast, slow in (("sma", pqo.sma, 5, 10), ("sma", pqo.sma, 10, 20), ("ema", pqo.ema, 12, 26), ("sma", pqo.sma, 20, 7)):
        sig = panel.ma_cross(fast, slow, ma_type).copy()
        assert set(np.unique(sig).tolist()) <= {-1, 0, 1} and (sig != 0).sum() > 1000
        for s in (0, 4, 33, 50, 69):
            a = int(starts[s])
            f, g = fn(d["close"][s][a:], fast), fn(d["close"][s][a:], slow)
            fv = np.full(N, np.nan); gv = np.full(N, np.nan)
            fv[a:] = np.where(f[1], f[0], np.nan); gv[a:] = np.where(g[1], g[0], np.nan)
            assert np.array_equal(sig[s], cross(fv, gv)), (ma_type, fast, slow, s)
    panel.close()
