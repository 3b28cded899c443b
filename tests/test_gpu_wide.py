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
