"""Many rolling windows in one launch (BASELINE config 5) against the oracle: every value and validity bit identical to
the per-window reference calls (STOCH / KDJ momentum.py:178-186 + D3, willr momentum.rs:630, midprice overlap.rs:281,
Donchian lines D3, atr volatility.rs:18)."""
import numpy as np
import pytest

import devread
import synth
import tolerances as T
from oracle import pqo

pytestmark = pytest.mark.gpu


def _refs(h, l, c, kdj, ext, atr):
    out = {}
    for k in kdj:
        K, D, J = pqo.kdj(h, l, c, k, 3, 3)
        out["kdj_k_%d" % k], out["kdj_d_%d" % k], out["kdj_j_%d" % k] = K, D, J
    for p in ext:
        out["willr_%d" % p] = pqo.willr(h, l, c, p)
        out["midprice_%d" % p] = pqo.midprice(h, l, p)
        up, lo = pqo.donchian(h, l, p)
        out["donchian_upper_%d" % p], out["donchian_lower_%d" % p] = up, lo
    if atr:
        out["atr_%d" % atr] = pqo.atr(h, l, c, atr)
    return out


def test_window_suite_small_panels_and_listing_dates():
    from polars_quant_b200.windows import WindowPanel
    S, N = 75, 900
    d = synth.ohlcv(S, N, seed=21)
    kdj, ext, atr = (5, 9, 14, 60, 250), (5, 20, 55, 250), 14
    wp = WindowPanel(S, N, kdj=kdj, ext=ext, atr=atr)
    starts = np.zeros(S, dtype=np.int32)
    starts[[3, 40, 74]] = (17, 300, 640)
    wp.panel.set_fields(close=d["close"], high=d["high"], low=d["low"], starts=starts)
    res = wp.compute()
    assert len(res) == 15 + 16 + 1
    for s in (0, 3, 31, 32, 40, 74):
        a = int(starts[s])
        refs = _refs(d["high"][s, a:], d["low"][s, a:], d["close"][s, a:], kdj, ext, atr)
        for name, (rv, rk) in refs.items():
            fv = np.full(N, np.nan); fk = np.zeros(N, bool)
            fv[a:], fk[a:] = rv, rk
            nbad, msg = T.compare(name, res[name][0][s], res[name][1][s], fv, fk)
            assert nbad == 0, f"symbol {s}: {msg}"
    wp.close()
    # other shapes of the unit dealing: one window, no ATR, windows on both sides of the shared / global limit
    for kdj, ext, atr in (((9,), (), 0), ((), (64, 65), 7), ((3, 100), (2,), 0), ((33, 129), (40, 128, 499), 3), ((600,), (32, 33, 501), 0)):
        wp = WindowPanel(40, 500, kdj=kdj, ext=ext, atr=atr)
        wp.panel.set_fields(close=d["close"][:40, :500], high=d["high"][:40, :500], low=d["low"][:40, :500])
        res = wp.compute()
        for s in (0, 39):
            for name, (rv, rk) in _refs(d["high"][s, :500], d["low"][s, :500], d["close"][s, :500], kdj, ext, atr).items():
                nbad, msg = T.compare(name, res[name][0][s], res[name][1][s], rv, rk)
                assert nbad == 0, f"{kdj} {ext} {atr} symbol {s}: {msg}"
        wp.close()


def test_config5_full_shape_against_the_oracle():
    """10,000 x 5,040, KDJ(5, 9, 14, 60, 250) + WILLR / MIDPRICE / Donchian(5, 20, 55, 250) + ATR(14), device-resident:
    symbol blocks across the grid are read back from HBM and compared with the oracle bit for bit."""
    import polars_quant_b200 as pq
    from polars_quant_b200.windows import WindowPanel
    S, N = 10_000, 5_040
    lib = pq._native.lib()
    kdj, ext, atr = (5, 9, 14, 60, 250), (5, 20, 55, 250), 14
    wp = WindowPanel(S, N, kdj=kdj, ext=ext, atr=atr, host_staging=False)
    wp.fill_synthetic(seed=55, sigma=0.02)
    wp.run()
    wp.panel.sync()
    nb, bp = wp.panel.tiled_shape()
    h_ = wp.panel._h
    names = wp.names()
    for b in (0, 1, 147, 148, 200, nb - 1):
        ns = min(32, S - b * 32)
        c, h, l = (devread.read_block(lib.pqb_panel_device_field(h_, f), b, bp, N)[:ns] for f in (0, 1, 2))
        outs = {k: devread.read_block(lib.pqb_panel_device_output(h_, k), b, bp, N)[:ns] for k in names}
        oks = {k: devread.read_validity_rows(lib.pqb_panel_device_validity(h_, k), b * 32, ns, wp.panel.validity_pitch, N) for k in names}
        for s in (0, ns - 1):
            refs = _refs(h[s], l[s], c[s], kdj, ext, atr)
            for k, name in names.items():
                nbad, msg = T.compare(name, outs[k][s], oks[k][s], refs[name][0], refs[name][1])
                assert nbad == 0, f"block {b} symbol {s}: {msg}"
    wp.close()


def test_long_windows_ignore_nan_highs_and_lows_like_the_reference():
    """The two-level window extremes (windows > 32) fold with the reference's f64::max / f64::min semantics (momentum.rs:644-650:
    a NaN value is ignored), exactly like the one-level ones: willr / Donchian / midprice with NaN highs and lows."""
    from polars_quant_b200.windows import WindowPanel
    S, N = 33, 700
    d = synth.ohlcv(S, N, seed=77)
    h, l, c = d["high"].copy(), d["low"].copy(), d["close"].copy()
    rng = np.random.default_rng(5)
    for s in range(S):
        idx = rng.choice(N, size=25, replace=False)
        h[s, idx[:12]] = np.nan
        l[s, idx[12:]] = np.nan
    h[7, 100:420] = np.nan                      # longer than every window but one: whole windows without a comparable value
    kdj, ext, atr = (), (20, 60, 250), 0
    wp = WindowPanel(S, N, kdj=kdj, ext=ext, atr=atr)
    wp.panel.set_fields(close=c, high=h, low=l)
    res = wp.compute()
    for s in (0, 7, 32):
        for p in ext:
            rv, rk = pqo.willr(h[s], l[s], c[s], p)
            nbad, msg = T.compare("willr_%d" % p, res["willr_%d" % p][0][s], res["willr_%d" % p][1][s], rv, rk)
            assert nbad == 0, f"symbol {s}: {msg}"
    wp.close()
