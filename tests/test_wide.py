"""The wide `{symbol}_{column}` panel format (README.md:88-161 `load`): one file per symbol, full join on `date`,
`date` + `{symbol}_{column}` columns -- CPU side (file reading and the join; the GPU side is tests/test_gpu_wide.py)."""
import numpy as np
import pyarrow as pa
import pyarrow.csv as pcsv
import pyarrow.parquet as pq
import pytest

import synth
from polars_quant_b200 import wide


def _symbol_table(seed, dates):
    d = synth.ohlcv(1, len(dates), seed=seed)
    return pa.table({"date": pa.array(dates, type=pa.int32()), "open": d["open"][0], "high": d["high"][0], "low": d["low"][0],
                     "close": d["close"][0], "volume": d["volume"][0].astype(np.int64)})


@pytest.fixture()
def folder(tmp_path):
    pq.write_table(_symbol_table(1, list(range(100, 160))), tmp_path / "AAPL.parquet")
    pq.write_table(_symbol_table(2, list(range(120, 170))), tmp_path / "SH600000.parquet")       # listed later, ends later
    pcsv.write_csv(_symbol_table(3, [d for d in range(100, 160) if d not in (130, 131)]), tmp_path / "MSFT_daily.csv")   # a halt
    (tmp_path / "notes.txt").write_text("ignored")
    return tmp_path


def test_load_full_joins_on_date_and_names_columns_like_the_reference(folder):
    t = wide.load(folder)
    assert t.column_names[0] == "date"
    assert t["date"].to_pylist() == list(range(100, 170))                       # union of the dates, ascending
    for sym in ("AAPL", "SH600000", "MSFT_daily"):
        for f in ("open", "high", "low", "close", "volume"):
            assert f"{sym}_{f}" in t.column_names
    assert t["SH600000_close"].null_count == 20 and not t["SH600000_close"][0].is_valid and t["SH600000_close"][20].is_valid
    assert t["AAPL_close"].null_count == 10 and not t["AAPL_close"][69].is_valid          # delisted: trailing nulls
    assert t["MSFT_daily_close"].null_count == 12 and not t["MSFT_daily_close"][30].is_valid   # halt + the tail
    ref = _symbol_table(1, list(range(100, 160)))
    assert t["AAPL_high"].to_pylist()[:60] == ref["high"].to_pylist()


def test_load_filters(folder):
    assert set(c.split("_")[0] for c in wide.load(folder, file_type=["parquet"]).column_names[1:]) == {"AAPL", "SH600000"}
    assert all(c.startswith("SH600000_") for c in wide.load(folder, prefix="SH").column_names[1:])
    assert all(c.startswith("MSFT_daily_") for c in wide.load(folder, suffix="_daily").column_names[1:])
    with pytest.raises(FileNotFoundError):
        wide.load(folder, prefix="ZZ")
    with pytest.raises(NotImplementedError):
        wide.load(folder, file_type=["xlsx"])


def test_split_columns(folder):
    t = wide.load(folder)
    symbols, cols = wide.split_columns(t, wide.SUITE_FIELDS)
    assert symbols == ["AAPL", "MSFT_daily", "SH600000"]
    assert cols["close"]["MSFT_daily"] == "MSFT_daily_close" and set(cols["volume"]) == set(symbols)


def _expect_long(files, strategy, default):
    """Plain-Python reading of sequential.py:7-93: grid of dates x symbols, fill per symbol, then the default."""
    dates = sorted({d for rows in files.values() for d in rows})
    syms = sorted(files)
    out = {}
    for s in syms:
        col = [files[s].get(d) for d in dates]
        if strategy == "forward":
            last = None
            for i, v in enumerate(col):
                last = v if v is not None else last
                col[i] = last
        elif strategy == "backward":
            nxt = None
            for i in range(len(col) - 1, -1, -1):
                nxt = col[i] if col[i] is not None else nxt
                col[i] = nxt
        elif strategy == "zero":
            col = [0.0 if v is None else v for v in col]
        out[s] = [default if v is None else v for v in col]
    return dates, syms, out


@pytest.mark.parametrize("strategy", ["forward", "backward", "zero", "none"])
def test_prepare_sequential_data_aligns_fills_and_sorts(tmp_path, strategy):
    files = {"AAA": {d: 10.0 + d for d in range(5, 15)},                        # stops early
             "BBB": {d: 20.0 + d for d in range(8, 20) if d not in (11, 12)},   # listed later, a halt
             "CCC": {d: 30.0 + d for d in (3, 19)}}
    pq.write_table(pa.table({"date": pa.array(list(files["AAA"]), type=pa.int32()), "close": list(files["AAA"].values())}),
                   tmp_path / "AAA.parquet")
    pcsv.write_csv(pa.table({"date": pa.array(list(files["BBB"]), type=pa.int32()), "close": list(files["BBB"].values())}),
                   tmp_path / "BBB.csv")
    pq.write_table(pa.table({"date": pa.array(list(files["CCC"]), type=pa.int32()), "symbol": ["CCC", "CCC"],
                             "close": list(files["CCC"].values())}), tmp_path / "whatever.pqt")   # carries its own symbol column
    (tmp_path / "notes.txt").write_text("ignored")
    t = wide.prepare_sequential_data(tmp_path, fill_null_strategy=strategy, default_fill_value=-1.0)
    dates, syms, want = _expect_long(files, strategy, -1.0)
    assert t.column_names == ["date", "symbol", "close"] and t.num_rows == len(dates) * len(syms)
    assert t["date"].to_pylist() == [d for d in dates for _ in syms]            # sorted by (date, symbol)
    assert t["symbol"].to_pylist() == syms * len(dates)
    got = np.asarray(t["close"].to_numpy()).reshape(len(dates), len(syms))
    for k, s in enumerate(syms):
        assert got[:, k].tolist() == want[s], (s, strategy)
    assert t["close"].null_count == 0
    w = wide.to_wide(t)
    assert w.column_names == ["date", "AAA_close", "BBB_close", "CCC_close"] and w["date"].to_pylist() == dates
    assert w["BBB_close"].to_pylist() == want["BBB"]


def test_prepare_sequential_data_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        wide.prepare_sequential_data(tmp_path / "missing")
    (tmp_path / "x.txt").write_text("no data")
    with pytest.raises(ValueError):
        wide.prepare_sequential_data(tmp_path)
    pq.write_table(pa.table({"date": pa.array([1, 1], type=pa.int32()), "close": [1.0, 2.0]}), tmp_path / "AAA.parquet")
    with pytest.raises(ValueError, match="more than once"):
        wide.prepare_sequential_data(tmp_path)


def test_to_wide_leaves_absent_pairs_null():
    t = pa.table({"date": pa.array([1, 1, 2, 3], type=pa.int32()), "symbol": ["B", "A", "A", "B"], "close": [1.0, 2.0, 3.0, 4.0],
                  "volume": pa.array([10, 20, 30, 40], type=pa.int64())})
    w = wide.to_wide(t)
    assert w.column_names == ["date", "B_close", "B_volume", "A_close", "A_volume"]
    assert w["B_close"].to_pylist() == [1.0, None, 4.0] and w["A_volume"].to_pylist() == [20, 30, None]
