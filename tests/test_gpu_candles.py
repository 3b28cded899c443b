"""Candle engine on the GPU against the golden vectors made from the reference's own pattern.rs text:
all 61 Int32 pattern columns, the price transforms and BOP, bit-exact, through the panel path, the
single-column path and with the reference's null rules."""
from pathlib import Path

import numpy as np
import pytest

import polars_quant_b200 as pq
from polars_quant_b200 import _native as N
from polars_quant_b200 import candles

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def g():
    return np.load(ROOT / "tests" / "golden" / "candle_golden.npz")


def _panel(g, **kw):
    o, h, l, c = (g[k] for k in ("open", "high", "low", "close"))
    p = candles.CandlePanel(o.shape[0], o.shape[1], **kw)
    p.set_fields(o, h, l, c)
    return p


def test_all_61_patterns_match_the_reference_vectors(g):
    p = _panel(g)
    p.run_host()
    names = [str(n) for n in g["names"]]
    bad = []
    for k, n in enumerate(names):
        got = p.pattern(k)
        if not np.array_equal(got, g["patterns"][k].astype(np.int32)):
            i = np.argwhere(got != g["patterns"][k])[0]
            bad.append((n, int((got != g["patterns"][k]).sum()), i.tolist()))
    assert not bad, bad
    p.close()


def test_penetration_parameter(g):
    p = _panel(g)
    p.run_host(candles.default_params(penetration=float(g["penetration_value"])))
    names = [str(n) for n in g["names"]]
    for j, n in enumerate(str(x) for x in g["penetration_names"]):
        assert np.array_equal(p.pattern(names.index(n)), g["patterns_pen"][j].astype(np.int32)), n
    p.close()


def test_price_transforms_and_bop_are_bit_exact(g):
    p = _panel(g)
    p.run_host()
    for k, n in enumerate(N.PRICE_NAMES):
        v, ok = p.price(k)
        assert ok.all(), n
        assert np.array_equal(v.view(np.uint64), g["prices"][k].view(np.uint64)), n
    p.close()


def test_partial_masks_and_device_resident_run_agree(g):
    names = [str(n) for n in g["names"]]
    want = [names.index(n) for n in ("cdlengulfing", "cdlhammer", "cdlrisefall3methods")]
    mask = sum(1 << k for k in want)
    p = _panel(g, patterns_mask=mask, prices_mask=1 << 4)
    p.run_host()
    for k in want:
        assert np.array_equal(p.pattern(k), g["patterns"][k].astype(np.int32))
    with pytest.raises(ValueError):
        p.pattern(0)
    with pytest.raises(N.PqbError):
        p.run_host(candles.default_params())            # asks for planes this panel does not have
    p.close()


def test_single_column_entry_point_and_ragged_lengths(g):
    names = [str(n) for n in g["names"]]
    o, h, l, c = (g[k] for k in ("open", "high", "low", "close"))
    for n in ("cdl3outside", "cdlbreakaway", "cdldoji", "cdlmorningstar"):
        k = names.index(n)
        for s in (0, 17, o.shape[0] - 1):
            assert np.array_equal(candles.cdl(n, o[s], h[s], l[s], c[s]), g["patterns"][k][s]), (n, s)
    # a column shorter than the lookback / not a multiple of the tile: prefix of the same series
    k = names.index("cdlengulfing")
    for m in (1, 2, 5, 33, 257, 300):
        got = candles.cdl(k, o[3][:m], h[3][:m], l[3][:m], c[3][:m])
        assert np.array_equal(got, g["patterns"][k][3][:m]), m      # patterns only look back
    k = names.index("cdlpiercing")
    j = [str(x) for x in g["penetration_names"]].index("cdlpiercing")
    assert np.array_equal(candles.cdl(k, o[5], h[5], l[5], c[5], penetration=float(g["penetration_value"])), g["patterns_pen"][j][5])


def test_null_rules(g):
    o, h, l, c = (g[k][:8] for k in ("open", "high", "low", "close"))
    ok_h = np.ones(h.shape, dtype=bool)
    ok_h[2, 10:13] = False
    ok_c = np.ones(c.shape, dtype=bool)
    ok_c[5, 0] = False
    p = candles.CandlePanel(8, o.shape[1], patterns_mask=0, prices_mask=0b01111)
    p.set_fields(o, h, l, c, validity={1: ok_h, 3: ok_c})
    p.run_host(candles.default_params(patterns=0, prices=0b01111))
    avg, ok = p.price(0)
    assert np.array_equal(ok, ok_h & ok_c)                              # price.rs:24-27 all four valid
    med, okm = p.price(1)
    assert np.array_equal(okm, ok_h)                                    # :44-47 high and low only
    assert np.isnan(avg[~ok]).all() and np.array_equal(avg[ok], ((o + h + l + c) * 0.25)[ok])
    p.close()
    q = candles.CandlePanel(8, o.shape[1])
    q.set_fields(o, h, l, c, validity={1: ok_h})
    with pytest.raises(N.PqbError, match="not contiguous"):            # cdl* / bop: cont_slice()? pattern.rs:12
        q.run_host()
    q.close()


def test_large_panel_properties():
    """BASELINE config-2 shape on the device-resident path: determinism, value set, lookback zeros,
    doji consistency between the flag-based patterns."""
    S, n = 5000, 2520
    names = candles.pattern_names()
    ks = [names.index(x) for x in ("cdldoji", "cdllongleggeddoji", "cdlbreakaway", "cdlmarubozu", "cdlclosingmarubozu")]
    mask = sum(1 << k for k in ks)
    p = candles.CandlePanel(S, n, patterns_mask=mask, prices_mask=0b00011)
    p.fill_synthetic(seed=9, to_host=True)
    prm = candles.default_params(patterns=mask, prices=0b00011)
    p.run_host(prm)
    a = {k: p.pattern(k).copy() for k in ks}
    p.run_host(prm)
    for k in ks:
        assert np.array_equal(a[k], p.pattern(k))
        assert set(np.unique(a[k]).tolist()) <= {-100, 0, 100}
    assert (a[names.index("cdlbreakaway")][:, :4] == 0).all()
    assert (a[names.index("cdldoji")] != 0).sum() > 1000
    assert ((a[names.index("cdllongleggeddoji")] != 0) <= (a[names.index("cdldoji")] != 0)).all()
    assert ((a[names.index("cdlmarubozu")] != 0) <= (a[names.index("cdlclosingmarubozu")] != 0)).all()
    o, h, l, c = (p.host_field(f) for f in range(4))
    avg, ok = p.price(0)
    assert ok.all() and np.array_equal(avg, (o + h + l + c) * 0.25)
    p.close()


def test_candle_functions_through_the_polars_plugin_symbols(g):
    import pyarrow as pa
    from polars_quant_b200 import plugin
    names = [str(n) for n in g["names"]]
    s = 11
    cols = [pa.array(g[k][s]) for k in ("open", "high", "low", "close")]
    for n in ("cdl3inside", "cdlengulfing", "cdlkicking", "cdlxsidegap3methods", "cdlshortline"):
        out = plugin.call(n, cols)
        assert out.type == pa.int32() and out.null_count == 0
        assert np.array_equal(out.to_numpy(), g["patterns"][names.index(n)][s]), n
    j = [str(x) for x in g["penetration_names"]].index("cdleveningstar")
    out = plugin.call("cdleveningstar", cols + [float(g["penetration_value"])])
    assert np.array_equal(out.to_numpy(), g["patterns_pen"][j][s])
    o, h, l, c = cols
    for n, args, k in (("avgprice", [o, h, l, c], 0), ("medprice", [h, l], 1), ("typprice", [h, l, c], 2),
                       ("wclprice", [h, l, c], 3), ("bop", [o, h, l, c], 4)):
        out = plugin.call(n, args)
        assert out.null_count == 0 and np.array_equal(out.to_numpy().view(np.uint64), g["prices"][k][s].view(np.uint64)), n
    hn = pa.array(g["high"][s], mask=np.arange(g["high"].shape[1]) == 7)
    out = plugin.call("typprice", [hn, l, c])
    assert out.null_count == 1 and not out[7].is_valid                  # price.rs:66-69 null-propagating
    with pytest.raises(plugin.PluginError, match="not contiguous"):
        plugin.call("cdldoji", [o, hn, l, c])                           # pattern.rs:14 cont_slice()?
    with pytest.raises(plugin.PluginError, match="not contiguous"):
        plugin.call("bop", [o, hn, l, c])                               # momentum.rs:119


def test_every_pattern_through_the_reference_shim_names(g):
    """polars_quant_b200.talib carries the reference's 61 CDL* names (+ AVGPRICE / MEDPRICE / TYPPRICE / WCLPRICE / BOP) with the
    reference's signatures: every one of them, called like the reference's shim, gives the golden column."""
    import pyarrow as pa
    from polars_quant_b200 import talib
    names = [str(n) for n in g["names"]]
    pen_names = [str(x) for x in g["penetration_names"]]
    s = 7
    o, h, l, c = (pa.array(g[k][s]) for k in ("open", "high", "low", "close"))
    for k, n in enumerate(names):
        f = getattr(talib, n.upper())
        out = f(o, h, l, c)
        assert out.type == pa.int32(), n
        if n in pen_names:
            # (the golden default columns use the Rust side's default when no fifth input arrives; the Python shim always sends
            # ITS default, so the default call is checked against the engine's own single-column entry at that value)
            from polars_quant_b200 import candles
            import inspect
            d = inspect.signature(f).parameters["penetration"].default
            assert np.array_equal(out.to_numpy(), candles.cdl(k, g["open"][s], g["high"][s], g["low"][s], g["close"][s], penetration=d)), n
            out = f(o, h, l, c, penetration=float(g["penetration_value"]))
            assert np.array_equal(out.to_numpy(), g["patterns_pen"][pen_names.index(n)][s]), n
        else:
            assert np.array_equal(out.to_numpy(), g["patterns"][k][s]), n
    for k, out in enumerate((talib.AVGPRICE(o, h, l, c), talib.MEDPRICE(h, l), talib.TYPPRICE(h, l, c), talib.WCLPRICE(h, l, c),
                             talib.BOP(o, h, l, c))):
        assert np.array_equal(out.to_numpy().view(np.uint64), g["prices"][k][s].view(np.uint64)), k


def test_config2_size_panels_against_the_c_oracle():
    """Beyond the golden panel: full BASELINE config-2 shape (5,000 x 2,520) of random-walk OHLC and of the busy
    synthetic candles, every pattern column and price output against oracle/pq_candles.c, bit for bit."""
    from oracle import pqo
    S, n = 5000, 2520
    p = candles.CandlePanel(S, n)
    for fill in ("walk", "busy"):
        if fill == "walk":
            p.fill_random_walk(seed=0xC0FFEE, sigma=0.02, to_host=True)
        else:
            p.fill_synthetic(seed=17, to_host=True)
        p.run_host()
        o, h, l, c = (np.ascontiguousarray(p.host_field(f)) for f in range(4))
        pat, pr, _ = pqo.candles_panel(o, h, l, c)
        hits = 0
        for k in range(N.N_PATTERNS):
            got = p.pattern(k)
            assert np.array_equal(got, pat[k]), (fill, candles.pattern_names()[k], int((got != pat[k]).sum()))
            hits += int((got != 0).sum())
        assert hits > 1_000_000
        for k in range(N.N_PRICES):
            v, ok = p.price(k)
            assert ok.all() and np.array_equal(v.view(np.uint64), pr[k].view(np.uint64)), (fill, N.PRICE_NAMES[k])
    p.close()
