"""Candle engine, CPU side: the committed golden vectors (tests/golden/candle_golden.npz, made by executing
the reference's pattern.rs text -- tests/golden/make_pattern_golden.py) are complete and self-consistent, the
library's pattern table is the reference's order of definition, and the candle entry points fail loudly
without a device."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def g():
    return np.load(ROOT / "tests" / "golden" / "candle_golden.npz")


def test_golden_covers_all_61_patterns_and_every_one_that_can_fire(g):
    names = [str(n) for n in g["names"]]
    assert len(names) == 61 and names == sorted(names) and names[0] == "cdl2crows" and names[-1] == "cdlxsidegap3methods"
    pat = g["patterns"]
    assert pat.shape[0] == 61 and pat.shape[1:] == g["open"].shape
    assert set(np.unique(pat).tolist()) <= {-100, 0, 100}
    fired = {n: int((pat[k] != 0).sum()) for k, n in enumerate(names)}
    # cdl2crows cannot fire in the reference (pattern.rs:31 contradicts :27); everything else must be exercised
    assert [n for n, c in fired.items() if c == 0] == ["cdl2crows"]
    both = {"cdl3inside", "cdl3outside", "cdlengulfing", "cdlharami", "cdlbelthold", "cdlmarubozu", "cdlkicking"}
    for n in both:
        k = names.index(n)
        assert (pat[k] > 0).any() and (pat[k] < 0).any(), n
    assert [str(n) for n in g["penetration_names"]] == ["cdldarkcloudcover", "cdleveningdojistar", "cdleveningstar",
                                                         "cdlmorningdojistar", "cdlmorningstar", "cdlpiercing"]
    k = names.index("cdlpiercing")
    assert (g["patterns_pen"][5] != pat[k]).any()            # the second penetration value changes the outcome


def test_golden_prices_follow_price_rs(g):
    o, h, l, c = (g[k] for k in ("open", "high", "low", "close"))
    p = g["prices"]
    assert np.array_equal(p[0], (o + h + l + c) * 0.25)       # price.rs:25
    assert np.array_equal(p[1], (h + l) * 0.5)                # :45
    assert np.array_equal(p[2], (h + l + c) / 3.0)            # :67
    assert np.array_equal(p[3], (h + l + 2.0 * c) / 4.0)      # :89
    d = h - l
    with np.errstate(invalid="ignore", divide="ignore"):
        assert np.array_equal(p[4], np.where(d == 0.0, 0.0, (c - o) / np.where(d == 0.0, 1.0, d)))   # momentum.rs:130


def test_library_pattern_table_is_the_reference_order(g):
    from polars_quant_b200 import _native, candles
    _native.build()
    assert candles.pattern_names() == [str(n) for n in g["names"]]
    L = _native.lib()
    assert L.pqb_pattern_index(b"cdlhammer") == candles.pattern_names().index("cdlhammer")
    assert L.pqb_pattern_index(b"nope") == -1 and L.pqb_pattern_name(61) is None
    p = candles.default_params()
    assert p.patterns == (1 << 61) - 1 and p.prices == 31 and p.pen_piercing == 0.3 and p.pen_darkcloudcover == 0.3


def test_candle_entry_points_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from polars_quant_b200 import _native as N
    h = C.c_void_p()
    assert N.lib().pqb_candles_create(None, 4, 100, 1, 1, 1, C.byref(h)) == -3 and not h.value
    x = np.arange(8.0)
    col = N.Col(x.ctypes.data_as(C.c_void_p), None, 0, 8)
    out = np.zeros(8, dtype=np.int32)
    rc = N.lib().pqb_cdl(None, 3, C.byref(col), C.byref(col), C.byref(col), C.byref(col), 0.3, out.ctypes.data_as(C.c_void_p))
    assert rc == -3


def test_c_oracle_reproduces_every_golden_column(g):
    """oracle/pq_candles.c (the reference's per-function loops restated in C) against the vectors made by executing the
    reference's own text: all 61 patterns, both penetration settings, the price transforms and BOP, exactly."""
    from oracle import pqo
    o, h, l, c = (g[k] for k in ("open", "high", "low", "close"))
    pat, pr, used = pqo.candles_panel(o, h, l, c)
    names = [str(n) for n in g["names"]]
    assert used >= 1
    for k, n in enumerate(names):
        assert np.array_equal(pat[k], g["patterns"][k].astype(np.int32)), n
    for k in range(5):
        assert np.array_equal(pr[k].view(np.uint64), g["prices"][k].view(np.uint64)), k
    pv = float(g["penetration_value"])
    for j, n in enumerate(str(x) for x in g["penetration_names"]):
        k = names.index(n)
        for s in range(0, o.shape[0], 7):
            assert np.array_equal(pqo.cdl(k, o[s], h[s], l[s], c[s], pv), g["patterns_pen"][j][s]), n
