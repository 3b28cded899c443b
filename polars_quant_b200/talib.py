"""Python-level mirror of the reference's `polars_quant.talib` shims for the hot-path functions
(python/polars_quant/talib/{overlap,momentum,volatility,volume}.py): same names, same positional
parameters, same defaults, same tuple returns for the struct-valued functions.

Every function goes through libpqb200.so's polars plugin symbols (`_polars_plugin_<name>`):
* polars expressions / Series (when polars is installed): `register_plugin_function(plugin_path=
  libpqb200.so, function_name=<the reference's name>, args=[...], is_elementwise=False)` -- the
  reference's own call with only `plugin_path` changed;
* pyarrow arrays / chunked arrays / numpy arrays (this image has no polars): the same symbols called
  through the Arrow C Data Interface by `polars_quant_b200.plugin.call`, returning pyarrow arrays with
  Arrow nulls at the warm-up positions.
There is no CPU path: without a B200 the call raises PluginError("... PQB_ERR_NO_DEVICE ...").
"""
from __future__ import annotations

from . import plugin as _plugin

try:                                                      # not in this image; the branch is the reference's shim verbatim
    import polars as _pl
    from polars.plugins import register_plugin_function as _register
except Exception:                                         # pragma: no cover
    _pl = None
    _register = None

_LIB = _plugin.LIB_PATH


def _is_polars(x) -> bool:
    return _pl is not None and isinstance(x, (_pl.Expr, _pl.Series))


def _call(name, cols, params, fields=None):
    if _is_polars(cols[0]):
        expr = _register(args=[*cols, *params], plugin_path=_LIB, function_name=name, is_elementwise=False)
        if isinstance(cols[0], _pl.Series):
            if fields:
                df = cols[0].to_frame().select(expr).unnest(expr.meta.output_name())
                return tuple(df[c] for c in df.columns)
            return cols[0].to_frame().select(expr).to_series()
        return tuple(expr.struct.field(f) for f in fields) if fields else expr
    out = _plugin.call(name, [*cols, *params])
    if fields:
        return tuple(out.field(f) for f in fields)
    return out


# ---- overlap.py -------------------------------------------------------------------------------------
def SMA(real, timeperiod: int = 30):
    """SMA - Simple Moving Average (calc_sma overlap.rs:871)"""
    return _call("sma", [real], [timeperiod])


def EMA(real, timeperiod: int = 30):
    """EMA - Exponential Moving Average (calc_ema overlap.rs:660)"""
    return _call("ema", [real], [timeperiod])


def TEMA(real, timeperiod: int = 30):
    """TEMA - Triple Exponential Moving Average (calc_tema overlap.rs:1177)"""
    return _call("tema", [real], [timeperiod])


def TRIMA(real, timeperiod: int = 30):
    """TRIMA - Triangular Moving Average (calc_trima overlap.rs:1313)"""
    return _call("trima", [real], [timeperiod])


def MA(real, timeperiod: int = 30, matype: int = 0):
    """MA - Moving Average (calc_ma overlap.rs:857; matype 0/7 SMA, 1 EMA, 4 TEMA, 5 TRIMA)"""
    return _call("ma", [real], [timeperiod, matype])


def BBANDS(real, timeperiod: int = 20, nbdevup: float = 2.0, nbdevdn: float = 2.0):
    """BBANDS - Bollinger Bands (Upper, Middle, Lower) (overlap.rs:47)"""
    return _call("bbands", [real], [timeperiod, float(nbdevup), float(nbdevdn)], ("bb_upper", "bb_middle", "bb_lower"))


def MIDPOINT(real, timeperiod: int = 14):
    """MIDPOINT - MidPoint over period (overlap.rs:180, literal semantics)"""
    return _call("midpoint", [real], [timeperiod])


def MIDPRICE(high, low, timeperiod: int = 14):
    """MIDPRICE - Midpoint Price over period (overlap.rs:281)"""
    return _call("midprice", [high, low], [timeperiod])


# ---- momentum.py ------------------------------------------------------------------------------------
def RSI(real, timeperiod: int = 14):
    """RSI - Relative Strength Index (momentum.rs:507)"""
    return _call("rsi", [real], [timeperiod])


def MACD(real, fastperiod: int = 12, slowperiod: int = 26, signalperiod: int = 9):
    """MACD - Moving Average Convergence/Divergence (MACD, Signal, Hist) (momentum.rs:250)"""
    return _call("macd", [real], [fastperiod, slowperiod, signalperiod], ("macd", "macd_signal", "macd_hist"))


def MACDFIX(real, signalperiod: int = 9):
    """MACDFIX - Moving Average Convergence/Divergence Fixed 12/26/9"""
    return MACD(real, 12, 26, signalperiod)


def WILLR(high, low, close, timeperiod: int = 14):
    """WILLR - Williams' %R (momentum.rs:630)"""
    return _call("willr", [high, low, close], [timeperiod])


def MOM(real, timeperiod: int = 10):
    """MOM - Momentum (momentum.rs:384)"""
    return _call("mom", [real], [timeperiod])


def ROC(real, timeperiod: int = 10):
    """ROC - Rate of change : ((real/prevPrice)-1)*100 (momentum.rs:439)"""
    return _call("roc", [real], [timeperiod])


def ROCP(real, timeperiod: int = 10):
    """ROCP - Rate of change Percentage: (real-prevPrice)/prevPrice (momentum.rs:456)"""
    return _call("rocp", [real], [timeperiod])


def ROCR(real, timeperiod: int = 10):
    """ROCR - Rate of change ratio: (real/prevPrice) (momentum.rs:473)"""
    return _call("rocr", [real], [timeperiod])


def ROCR100(real, timeperiod: int = 10):
    """ROCR100 - Rate of change ratio 100 scale: (real/prevPrice)*100 (momentum.rs:490)"""
    return _call("rocr100", [real], [timeperiod])


def CMO(real, timeperiod: int = 14):
    """CMO - Chande Momentum Oscillator (momentum.rs:181)"""
    return _call("cmo", [real], [timeperiod])


def MFI(high, low, close, volume, timeperiod: int = 14):
    """MFI - Money Flow Index (momentum.rs:286)"""
    return _call("mfi", [high, low, close, volume], [timeperiod])


def CCI(high, low, close, timeperiod: int = 14):
    """CCI - Commodity Channel Index (momentum.rs:138)"""
    return _call("cci", [high, low, close], [timeperiod])


def ADX(high, low, close, timeperiod: int = 14):
    """ADX - Average Directional Movement Index (momentum.rs:11)"""
    return _call("adx", [high, low, close], [timeperiod])


def ADXR(high, low, close, timeperiod: int = 14):
    """ADXR - Average Directional Movement Index Rating (momentum.rs:29)"""
    return _call("adxr", [high, low, close], [timeperiod])


def DX(high, low, close, timeperiod: int = 14):
    """DX - Directional Movement Index (momentum.rs:226)"""
    return _call("dx", [high, low, close], [timeperiod])


def PLUS_DI(high, low, close, timeperiod: int = 14):
    """PLUS_DI - Plus Directional Indicator (momentum.rs:401; the reference returns calc_dm().0, i.e. DX)"""
    return _call("plus_di", [high, low, close], [timeperiod])


def MINUS_DI(high, low, close, timeperiod: int = 14):
    """MINUS_DI - Minus Directional Indicator (momentum.rs:346)"""
    return _call("minus_di", [high, low, close], [timeperiod])


def PLUS_DM(high, low, timeperiod: int = 14):
    """PLUS_DM - Plus Directional Movement (momentum.rs:418)"""
    return _call("plus_dm", [high, low], [timeperiod])


def MINUS_DM(high, low, timeperiod: int = 14):
    """MINUS_DM - Minus Directional Movement (momentum.rs:362)"""
    return _call("minus_dm", [high, low], [timeperiod])


def TRIX(real, timeperiod: int = 30):
    """TRIX - 1-day Rate-Of-Change (ROC) of a Triple Smooth EMA (momentum.rs:544)"""
    return _call("trix", [real], [timeperiod])


def ULTOSC(high, low, close, timeperiod1: int = 7, timeperiod2: int = 14, timeperiod3: int = 28):
    """ULTOSC - Ultimate Oscillator (momentum.rs:573)"""
    return _call("ultosc", [high, low, close], [timeperiod1, timeperiod2, timeperiod3])


def AROON(high, low, timeperiod: int = 14):
    """AROON - Aroon (aroon_up, aroon_down) (momentum.rs:63, momentum.py:32-38)"""
    return _call("aroon", [high, low], [timeperiod], ("aroon_up", "aroon_down"))


def STOCH(high, low, close, fastk_period: int = 5, slowk_period: int = 3, slowk_matype: int = 0,
          slowd_period: int = 3, slowd_matype: int = 0):
    """STOCH - Stochastic (SlowK, SlowD).  The reference composes it in Python from polars rolling min / max and two
    MA plugin calls (momentum.py:178-186); here it is one call (fused kernel for SMA smoothings, a chain of device
    passes for the other matypes)."""
    return _call("stoch", [high, low, close], [fastk_period, slowk_period, slowk_matype, slowd_period, slowd_matype],
                 ("slowk", "slowd"))


def STOCHF(high, low, close, fastk_period: int = 5, fastd_period: int = 3, fastd_matype: int = 0):
    """STOCHF - Stochastic Fast (FastK, FastD) (momentum.py:188-195)"""
    return _call("stochf", [high, low, close], [fastk_period, fastd_period, fastd_matype], ("fastk", "fastd"))


def STOCHRSI(real, timeperiod: int = 14, fastk_period: int = 5, fastd_period: int = 3, fastd_matype: int = 0):
    """STOCHRSI - Stochastic Relative Strength Index (FastK, FastD) (momentum.py:197-205)"""
    return _call("stochrsi", [real], [timeperiod, fastk_period, fastd_period, fastd_matype], ("fastk_rsi", "fastd_rsi"))


def MACDEXT(real, fastperiod: int = 12, fastmatype: int = 0, slowperiod: int = 26, slowmatype: int = 0,
            signalperiod: int = 9, signalmatype: int = 0):
    """MACDEXT - MACD with controllable MA type (macd_dif, macd_dea, macd_hist) (momentum.py:83-88)"""
    return _call("macdext", [real], [fastperiod, fastmatype, slowperiod, slowmatype, signalperiod, signalmatype],
                 ("macd_dif", "macd_dea", "macd_hist"))


def KDJ(high, low, close, fastk_period: int = 9, k_period: int = 3, d_period: int = 3):
    """KDJ (K, D, J) := STOCH(9, 3, 3) with J = 3K - 2D (README.md:720-722)"""
    return _call("kdj", [high, low, close], [fastk_period, k_period, d_period], ("k", "d", "j"))


# ---- volatility.py ----------------------------------------------------------------------------------
def TRANGE(high, low, close):
    """TRANGE - True Range (volatility.rs:51)"""
    return _call("trange", [high, low, close], [])


def ATR(high, low, close, timeperiod: int = 14):
    """ATR - Average True Range (volatility.rs:18)"""
    return _call("atr", [high, low, close], [timeperiod])


def NATR(high, low, close, timeperiod: int = 14):
    """NATR - Normalized Average True Range (volatility.rs:34)"""
    return _call("natr", [high, low, close], [timeperiod])


# ---- volume.py --------------------------------------------------------------------------------------
def OBV(real, volume):
    """OBV - On Balance Volume (volume.rs:70)"""
    return _call("obv", [real, volume], [])


def AD(high, low, close, volume):
    """AD - Chaikin A/D Line (volume.rs:19)"""
    return _call("ad", [high, low, close, volume], [])


def ADOSC(high, low, close, volume, fastperiod: int = 3, slowperiod: int = 10):
    """ADOSC - Chaikin A/D Oscillator (volume.rs:34)"""
    return _call("adosc", [high, low, close, volume], [fastperiod, slowperiod])


# ---- price.py, momentum.py BOP, pattern.py (the candle engine) ------------------------
def AVGPRICE(open, high, low, close):
    """AVGPRICE - Average Price (price.rs:10)"""
    return _call("avgprice", [open, high, low, close], [])


def MEDPRICE(high, low):
    """MEDPRICE - Median Price (price.rs)"""
    return _call("medprice", [high, low], [])


def TYPPRICE(high, low, close):
    """TYPPRICE - Typical Price (price.rs)"""
    return _call("typprice", [high, low, close], [])


def WCLPRICE(high, low, close):
    """WCLPRICE - Weighted Close Price (price.rs)"""
    return _call("wclprice", [high, low, close], [])


def BOP(open, high, low, close):
    """BOP - Balance Of Power (momentum.rs:113)"""
    return _call("bop", [open, high, low, close], [])


# the 61 candlestick patterns of pattern.rs:9-2065 (Int32 in {-100, 0, 100}); nine take a `penetration` with the reference's
# defaults (python/polars_quant/talib/pattern.py)
_CDL_PENETRATION = {"CDLABANDONEDBABY": 0.3, "CDLDARKCLOUDCOVER": 0.5, "CDLEVENINGDOJISTAR": 0.3, "CDLEVENINGSTAR": 0.3,
                    "CDLMATHOLD": 0.5, "CDLMORNINGDOJISTAR": 0.3, "CDLMORNINGSTAR": 0.3, "CDLPIERCING": 0.5, "CDLTHRUSTING": 0.3}


def _make_cdl(upper: str):
    lower = upper.lower()
    if upper in _CDL_PENETRATION:
        default = _CDL_PENETRATION[upper]

        def f(o, h, l, c, penetration: float = default):
            return _call(lower, [o, h, l, c], [float(penetration)])
    else:
        def f(o, h, l, c):
            return _call(lower, [o, h, l, c], [])
    f.__name__ = f.__qualname__ = upper
    f.__doc__ = "%s (pattern.rs `%s`; the fused candle kernel, bit-identical to the reference's loop)" % (upper, lower)
    return f


def _define_patterns():
    from .candles import pattern_names
    names = []
    for lower in pattern_names():
        upper = lower.upper()
        globals()[upper] = _make_cdl(upper)
        names.append(upper)
    return names


_CDL_NAMES = _define_patterns()


# ---- names the reference's shims define but this build does not serve ------------------------------------------------
def _not_built(name, why):
    def f(*args, **kwargs):
        raise NotImplementedError("%s is not built: %s" % (name, why))
    f.__name__ = f.__qualname__ = name
    f.__doc__ = "%s -- not built: %s" % (name, why)
    return f


for _n in ("WMA", "DEMA", "T3", "KAMA"):          # defective in the reference (restated in the oracle only)
    globals()[_n] = _not_built(_n, "defective in the reference (SURVEY.md 8a); MA(matype) refuses it with PQB_ERR_UNSUPPORTED too")
for _n in ("APO", "PPO", "AROONOSC"):             # registered by the reference's Python shims without a Rust symbol
    globals()[_n] = _not_built(_n, "the reference registers the name but has no plugin symbol for it (momentum.py:27,42,138)")


__all__ = ["SMA", "EMA", "TEMA", "TRIMA", "MA", "BBANDS", "MIDPOINT", "MIDPRICE", "RSI", "MACD", "MACDFIX", "WILLR",
           "MOM", "ROC", "ROCP", "ROCR", "ROCR100", "CMO", "MFI", "CCI", "STOCH", "STOCHF", "STOCHRSI", "MACDEXT", "KDJ", "TRANGE", "ATR", "NATR",
           "OBV", "AD", "ADOSC", "ADX", "ADXR", "DX", "PLUS_DI", "MINUS_DI", "PLUS_DM", "MINUS_DM", "TRIX", "ULTOSC",
           "AROON", "AVGPRICE", "MEDPRICE", "TYPPRICE", "WCLPRICE", "BOP", *_CDL_NAMES]
