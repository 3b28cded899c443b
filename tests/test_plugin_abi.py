"""The polars expression-plugin boundary (include/pqb200_polars_plugin.h) on a machine WITHOUT a GPU:
every declared `_polars_plugin_*` symbol is exported, the planning-time field functions return the
reference's output fields (Float64 named after the first input; `bbands` / `macd_res` structs with the
reference's field names), parameter intake and input validation work, the callee consumes its inputs,
failures leave return_value untouched and set the last-error message -- and compute calls fail loudly
with PQB_ERR_NO_DEVICE instead of falling back to a CPU path."""
import re
from pathlib import Path

import numpy as np
import pyarrow as pa
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def plugin():
    from polars_quant_b200 import _native
    _native.build()
    from polars_quant_b200 import plugin as P
    return P


def declared_plugins():
    text = (ROOT / "include" / "pqb200_polars_plugin.h").read_text()
    return re.findall(r"^PQB_POLARS_PLUGIN\((\w+)\)", text, flags=re.M)


REFERENCE_NAMES = ["sma", "ema", "tema", "trima", "ma", "bbands", "midpoint", "midprice", "rsi", "macd", "willr", "mom",
                   "roc", "rocp", "rocr", "rocr100", "cmo", "mfi", "cci", "trange", "atr", "natr", "obv", "ad", "adosc"]


def test_every_declared_plugin_symbol_is_exported(plugin):
    from polars_quant_b200 import _native as N
    names = declared_plugins()
    assert set(REFERENCE_NAMES) <= set(names) and {"stoch", "kdj"} <= set(names)
    L = N.lib()
    for n in names:
        assert hasattr(L, "_polars_plugin_" + n), n
        assert hasattr(L, "_polars_plugin_field_" + n), n
    assert hasattr(L, "_polars_plugin_get_last_error_message")
    assert plugin.version() == (0, 1)


def test_candle_plugin_symbols(plugin):
    from polars_quant_b200 import _native as N, candles
    L = N.lib()
    names = candles.pattern_names() + ["avgprice", "medprice", "typprice", "wclprice", "bop"]
    assert len(names) == 66
    for n in names:
        assert hasattr(L, "_polars_plugin_" + n) and hasattr(L, "_polars_plugin_field_" + n), n
    ins = [pa.field(x, pa.float64()) for x in ("o", "h", "l", "c")]
    f = plugin.output_field("cdlengulfing", ins)
    assert f.name == "o" and f.type == pa.int32()                      # #[polars_expr(output_type=Int32)] pattern.rs:9
    assert plugin.output_field("bop", ins).type == pa.float64()
    e = pa.array([], type=pa.float64())
    out = plugin.call("cdlhammer", [e, e, e, e])
    assert out.type == pa.int32() and len(out) == 0
    assert plugin.call("medprice", [e, e]).type == pa.float64()
    with pytest.raises(plugin.PluginError, match="expected 4 input columns"):
        plugin.call("cdldoji", [e, e])


def test_field_functions_return_the_reference_output_fields(plugin):
    f = plugin.output_field("ema", [pa.field("AAPL_close", pa.float64()), pa.field("literal", pa.int32())])
    assert f.name == "AAPL_close" and f.type == pa.float64()          # FieldsMapper::with_dtype(Float64)
    f = plugin.output_field("atr", [pa.field("h", pa.float32()), pa.field("l", pa.float32()), pa.field("c", pa.float32())])
    assert f.name == "h" and f.type == pa.float64()
    bb = plugin.output_field("bbands", [pa.field("x", pa.float64())])
    assert bb.name == "bbands"                                        # bbands_output overlap.rs:30-38
    assert [c.name for c in bb.type] == ["bb_upper", "bb_middle", "bb_lower"]
    assert all(c.type == pa.float64() for c in bb.type)
    m = plugin.output_field("macd", [pa.field("x", pa.float64())])
    assert m.name == "macd_res"                                       # macd_output momentum.rs:239-247
    assert [c.name for c in m.type] == ["macd", "macd_signal", "macd_hist"]
    k = plugin.output_field("kdj", [pa.field("h", pa.float64())])
    assert [c.name for c in k.type] == ["k", "d", "j"]


def test_empty_columns_need_no_device_and_keep_the_schema(plugin):
    e = pa.array([], type=pa.float64())
    out = plugin.call("sma", [e, 5])
    assert out.type == pa.float64() and len(out) == 0
    out = plugin.call("bbands", [e], kwargs={"timeperiod": 5, "nbdevup": 1.5, "nbdevdn": None})
    assert pa.types.is_struct(out.type) and len(out) == 0
    assert [out.type.field(i).name for i in range(3)] == ["bb_upper", "bb_middle", "bb_lower"]
    out = plugin.call("macd", [pa.chunked_array([e, e]), 12, 26, 9])
    assert len(out) == 0 and [out.type.field(i).name for i in range(3)] == ["macd", "macd_signal", "macd_hist"]


def test_argument_errors_are_reported_through_the_last_error_message(plugin):
    x = pa.array(np.arange(10.0))
    with pytest.raises(plugin.PluginError, match="expected 3 input columns"):
        plugin.call("atr", [x, x])
    with pytest.raises(plugin.PluginError, match="differ in length"):
        plugin.call("obv", [x, pa.array(np.arange(9.0))])
    with pytest.raises(plugin.PluginError, match="cannot be cast to Float64"):
        plugin.call("sma", [pa.array(["a", "b"]), 2])
    with pytest.raises(plugin.PluginError, match="unknown kwarg 'window'"):
        plugin.call("sma", [pa.array([], type=pa.float64())], kwargs={"window": 3})
    with pytest.raises(plugin.PluginError, match="non-negative integer"):
        plugin.call("ema", [pa.array([], type=pa.float64()), -3])
    with pytest.raises(plugin.PluginError, match="non-negative integer"):
        plugin.call("ema", [pa.array([], type=pa.float64()), 2.5])


def test_compute_fails_loudly_without_a_device(plugin):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    x = pa.array(np.linspace(1.0, 2.0, 64))
    with pytest.raises(plugin.PluginError, match="PQB_ERR_NO_DEVICE|no CPU fallback|no CUDA"):
        plugin.call("sma", [x, 5])                       # inputs are still consumed (call() asserts it)
    from polars_quant_b200 import talib
    with pytest.raises(plugin.PluginError):
        talib.MACD(x)


def test_talib_mirror_has_the_reference_signatures():
    import inspect
    from polars_quant_b200 import talib
    want = {"SMA": ["real", "timeperiod"], "EMA": ["real", "timeperiod"], "MA": ["real", "timeperiod", "matype"],
            "BBANDS": ["real", "timeperiod", "nbdevup", "nbdevdn"], "MIDPRICE": ["high", "low", "timeperiod"],
            "MACD": ["real", "fastperiod", "slowperiod", "signalperiod"], "RSI": ["real", "timeperiod"],
            "ATR": ["high", "low", "close", "timeperiod"], "OBV": ["real", "volume"],
            "ADOSC": ["high", "low", "close", "volume", "fastperiod", "slowperiod"],
            "STOCH": ["high", "low", "close", "fastk_period", "slowk_period", "slowk_matype", "slowd_period", "slowd_matype"]}
    for name, params in want.items():
        assert list(inspect.signature(getattr(talib, name)).parameters) == params
    d = {k: v.default for k, v in inspect.signature(talib.BBANDS).parameters.items() if k != "real"}
    assert d == {"timeperiod": 20, "nbdevup": 2.0, "nbdevdn": 2.0}             # overlap.py:9-11
    assert inspect.signature(talib.MOM).parameters["timeperiod"].default == 10   # momentum.py
