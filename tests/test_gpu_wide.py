"""Wide `{symbol}_{column}` tables through the GPU (polars_quant_b200.wide.WidePanel): symbols with different listing
dates (leading nulls), a delisting (trailing nulls) and a halt (interior nulls) against the oracle with the same
validity, bit for bit; zero-copy Arrow results; candles with the reference's null refusal."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import synth
import tolerances as T
from oracle import pqo
import polars_quant_b200 as pqb
from polars_quant_b200 import wide, _native as N
from test_wide import _symbol_table

pytestmark = pytest.mark.gpu


def _col(t, name):
    a = t[name].combine_chunks()
    ok = ~np.asarray(a.is_null())
    return np.where(ok, np.asarray(a.to_numpy(zero_copy_only=False), dtype=np.float64), np.nan), ok


def test_wide_suite_matches_the_oracle_per_symbol(tmp_path):
    pq.write_table(_symbol_table(1, list(range(0, 400))), tmp_path / "AAA.parquet")
    pq.write_table(_symbol_table(2, list(range(120, 430))), tmp_path / "BBB.parquet")                       # listed later
    pq.write_table(_symbol_table(3, [d for d in range(0, 430) if not 200 <= d < 204]), tmp_path / "CCC.parquet")   # a halt
    t = wide.load(tmp_path)
    wp = wide.WidePanel(t)
    names = ["sma", "ema", "bb_upper", "macd", "rsi", "atr", "obv", "ad", "kdj_k", "willr", "midprice"]
    out = wp.suite(outputs=names)
    assert out.column_names[0] == "date" and out.num_rows == 430 and out.num_columns == 1 + 3 * len(names)
    for sym in ("AAA", "BBB", "CCC"):
        c, okc = _col(t, sym + "_close")
        h, okh = _col(t, sym + "_high")
        l, okl = _col(t, sym + "_low")
        v, okv = _col(t, sym + "_volume")
        k = lambda m: None if m.all() else m.astype(np.uint8)
        c0, h0, l0, v0 = (np.where(m, x, 0.0) for x, m in ((c, okc), (h, okh), (l, okl), (v, okv)))
        refs = {"sma": pqo.sma(c0, 30, k(okc)), "ema": pqo.ema(c0, 30, k(okc)), "bb_upper": pqo.bbands(c0, 20, 2.0, 2.0, k(okc))[0],
                "atr": pqo.atr(h0, l0, c0, 14, k(okh), k(okl), k(okc)), "obv": pqo.obv(c0, v0, k(okc), k(okv)),
                "ad": pqo.ad(h0, l0, c0, v0, k(okh), k(okl), k(okc), k(okv))}
        if okc.all():                                   # momentum.rs functions: only null-free columns (AAA) are defined
            refs.update({"macd": pqo.macd(c0)[0], "rsi": pqo.rsi(c0, 14), "willr": pqo.willr(h0, l0, c0, 14),
                         "kdj_k": pqo.kdj(h0, l0, c0)[0], "midprice": pqo.midprice(h0, l0, 14)})
        for n, ref in refs.items():
            gv, gok = _col(out, f"{sym}_{n}")
            nbad, msg = T.compare(f"{sym}_{n}", gv, gok, ref[0], ref[1])
            assert nbad == 0, msg
    gv, gok = _col(out, "CCC_rsi")                      # interior nulls: rsi refuses the column -> all null for that symbol
    assert not gok.any()
    del wp                                              # the result aliases the panel's pinned memory and keeps it alive
    assert out["AAA_sma"].null_count >= 29 and np.isfinite(out["AAA_sma"][100].as_py())


def test_lazy_result_columns_are_the_table_columns(tmp_path):
    """`suite(lazy=True)`: columns wrapped on demand over the pinned planes are the columns of the full table, bit for bit and
    null for null; repeated calls reuse pooled planes (a result must keep ITS panel's memory, not the next call's)."""
    for k, days in enumerate((range(0, 300), range(40, 300), [d for d in range(0, 300) if d != 150])):
        pq.write_table(_symbol_table(10 + k, list(days)), tmp_path / ("S_%d.parquet" % k))          # (symbols with an underscore)
    t = wide.load(tmp_path)
    wp = wide.WidePanel(t)
    names = ["sma", "bb_upper", "macd_signal", "kdj_j", "obv"]
    full = wp.suite(outputs=names)
    lazy = wp.suite(outputs=names, lazy=True)
    again = wp.suite(outputs=names, lazy=True)           # a third panel while the first two results are alive
    assert lazy.symbols == ["S_0", "S_1", "S_2"] and sorted(lazy.outputs) == sorted(names) and len(lazy) == 15
    assert lazy.column_names == full.column_names[1:]
    for res in (lazy, again):
        for name in res.column_names:
            a, b = res[name], full[name].combine_chunks()
            assert a.null_count == b.null_count and a.equals(b), name
    v, ok = lazy.matrix("sma")
    assert v.shape == (3, 300) and np.array_equal(ok[0], ~np.asarray(full["S_0_sma"].combine_chunks().is_null()))
    assert set(lazy.symbol("S_1")) == set(names)
    assert lazy.table().equals(full)
    with pytest.raises(KeyError):
        lazy["S_0_rsi"]


def test_wide_suite_sharded_over_devices_equals_one_device(tmp_path):
    """`WidePanel.suite(devices=[...])`: the symbols sharded over GPUs (contiguous whole-block ranges, one host thread each; the
    same device twice on a one-GPU box) give the table of one device, eagerly and lazily."""
    n_dev = max(1, N.lib().pqb_device_count())
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    for k in range(70):                                        # three symbol blocks: shards of 32 + 38 symbols on two devices
        days = range(k % 5, 260) if k % 9 else [d for d in range(0, 260) if d != 120]
        pq.write_table(_symbol_table(100 + k, list(days)), tmp_path / ("T%03d.parquet" % k))
    t = wide.load(tmp_path)
    wp = wide.WidePanel(t)
    names = ["ema", "bb_lower", "atr", "kdj_d"]
    one = wp.suite(outputs=names)
    many = wp.suite(outputs=names, devices=devices)
    assert many.column_names == one.column_names and many.equals(one)
    lazy = wp.suite(outputs=names, devices=devices, lazy=True)
    assert lazy.column_names == one.column_names[1:]
    for name in ("T000_ema", "T031_atr", "T032_kdj_d", "T069_bb_lower"):
        assert lazy[name].equals(one[name].combine_chunks()), name
    v, ok = lazy.matrix("atr")
    assert v.shape == (70, 260) and np.array_equal(ok[40], ~np.asarray(one["T040_atr"].combine_chunks().is_null()))
    assert lazy.table().equals(one)


def test_wide_candles_and_null_refusal(tmp_path):
    pq.write_table(_symbol_table(5, list(range(0, 300))), tmp_path / "AAA.parquet")
    pq.write_table(_symbol_table(6, list(range(50, 300))), tmp_path / "BBB.parquet")
    t = wide.load(tmp_path)
    wp = wide.WidePanel(t)
    with pytest.raises(ValueError, match="BBB"):
        wp.candles()
    out = wp.candles(patterns=["cdlengulfing", "cdldoji"], prices=["typprice", "bop"], on_nulls="skip")
    assert out.column_names == ["date", "AAA_cdlengulfing", "AAA_cdldoji", "AAA_typprice", "AAA_bop"]
    o, h, l, c = (np.asarray(t["AAA_" + f].to_numpy()) for f in ("open", "high", "low", "close"))
    names = [n for n in __import__("polars_quant_b200.candles", fromlist=["x"]).pattern_names()]
    assert out["AAA_cdlengulfing"].type == pa.int32()
    assert np.array_equal(out["AAA_cdlengulfing"].to_numpy(), pqo.cdl(names.index("cdlengulfing"), o, h, l, c))
    assert np.array_equal(out["AAA_cdldoji"].to_numpy(), pqo.cdl(names.index("cdldoji"), o, h, l, c))
    assert np.array_equal(out["AAA_typprice"].to_numpy(), pqo.price(2, o, h, l, c))
    assert np.array_equal(out["AAA_bop"].to_numpy(), pqo.price(4, o, h, l, c))


def test_info_last_row_reductions_match_the_oracle(tmp_path):
    """Selector.info() columns (README-only in the reference; defined in include/pqb200.h): bit-exact against pqo_info,
    with symbols listed 3 / 10 / 25 bars before the end (statistics that need more history are null)."""
    S, NB = 70, 300
    d = synth.ohlcv(S, NB, seed=21)
    p = pqb.Panel(S, NB)
    starts = np.zeros(S, dtype=np.int32)
    starts[5], starts[33], starts[64], starts[69] = NB - 3, NB - 10, NB - 25, NB - 1
    for s in range(S):
        ok = np.arange(NB) >= starts[s]
        for f, name in enumerate(("close", "high", "low", "volume")):
            p.set_column(s, f, d[name][s], validity=None if starts[s] == 0 else np.packbits(ok, bitorder="little"))
    p.upload()
    got = p.info()
    for s in range(S):
        ref, rok = pqo.info(d["close"][s], d["high"][s], d["low"][s], d["volume"][s], int(starts[s]))
        for k, name in enumerate(N.INFO_NAMES):
            v, ok = got[name]
            assert ok[s] == rok[k], (s, name)
            if rok[k]:
                assert v[s].view(np.uint64) == ref[k].view(np.uint64), (s, name, v[s], ref[k])
            else:
                assert np.isnan(v[s])
    assert got["volatility"][1][64] and not got["volatility"][1][33] and got["ma_5"][1][33] and not got["ma_5"][1][5]
    # the wide-table route: one row per symbol, the README's 15 columns in its order
    pq.write_table(_symbol_table(1, list(range(0, 200))), tmp_path / "AAA.parquet")
    pq.write_table(_symbol_table(2, list(range(190, 200))), tmp_path / "BBB.parquet")
    t = wide.load(tmp_path)
    info = wide.WidePanel(t).info()
    assert info.column_names == ["symbol", "price", "open", "high", "low", "volume", "return_1d", "return_5d", "return_20d",
                                 "volatility", "ma_5", "ma_10", "ma_20", "volume_ratio", "amplitude"]
    assert info["symbol"].to_pylist() == ["AAA", "BBB"]
    assert info["price"][0].as_py() == t["AAA_close"][199].as_py() and info["open"][1].as_py() == t["BBB_open"][199].as_py()
    assert info["return_20d"][1].as_py() is None and info["return_5d"][1].as_py() is not None
    c = np.asarray(t["AAA_close"].to_numpy(), dtype=np.float64)
    assert info["ma_10"][0].as_py() == pytest.approx(c[-10:].mean(), rel=1e-14)
    r = c[-20:] / c[-21:-1] - 1.0
    assert info["volatility"][0].as_py() == pytest.approx(r.std(ddof=1) * np.sqrt(252.0) * 100.0, rel=1e-12)


def test_info_on_a_panel_shorter_than_every_window():
    """Three bars: the last-row columns and return_1d exist, everything that needs more history is null."""
    d = synth.ohlcv(5, 3, seed=8)
    p = pqb.Panel(5, 3)
    p.set_fields(d["close"], d["high"], d["low"], d["volume"])
    with pytest.raises(Exception):
        p.info()                                         # inputs not on the device yet
    p.upload()
    got = p.info()
    for s in range(5):
        ref, rok = pqo.info(d["close"][s], d["high"][s], d["low"][s], d["volume"][s])
        for k, name in enumerate(N.INFO_NAMES):
            assert got[name][1][s] == rok[k]
            if rok[k]:
                assert got[name][0][s].view(np.uint64) == ref[k].view(np.uint64)
    assert got["return_1d"][1].all() and not got["return_5d"][1].any() and not got["ma_5"][1].any() and got["amplitude"][1].all()
