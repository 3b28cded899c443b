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
