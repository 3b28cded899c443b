"""ctypes binding of the C oracle (oracle/pq_oracle.c).  TEST INFRASTRUCTURE ONLY.

Columns are (values: float64 ndarray, ok: bool ndarray | None).  Every wrapper returns
(values, ok) per output; null slots hold NaN.  Raises OracleError when the reference
itself would fail on the input (PQO_ERR_*).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libpq_oracle.so"

ERRORS = {-1: "nulls (cont_slice Err)", -2: "reference panics", -3: "shape Err", -4: "alloc"}


class OracleError(Exception):
    def __init__(self, code):
        super().__init__("oracle: reference would fail: %s" % ERRORS.get(code, code))
        self.code = code


_SO_CANDLES = _HERE / "libpq_candles.so"


def build(force: bool = False) -> Path:
    for so, src in ((_SO, _HERE / "pq_oracle.c"), (_SO_CANDLES, _HERE / "pq_candles.c")):
        if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
            env = dict(os.environ)
            env.setdefault("CC", "gcc")
            subprocess.run(["make", "-C", str(_HERE), "-B", so.name], check=True, env=env, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
    return _lib


class SuiteParams(C.Structure):
    """Mirror of pqo_suite_params (== the defaults of the reference's Python signatures)."""
    _fields_ = [("sma", C.c_int32), ("ema", C.c_int32), ("tema", C.c_int32), ("trima", C.c_int32),
                ("bb", C.c_int32), ("bb_up", C.c_double), ("bb_dn", C.c_double),
                ("macd_fast", C.c_int32), ("macd_slow", C.c_int32), ("macd_signal", C.c_int32),
                ("rsi", C.c_int32), ("atr", C.c_int32), ("natr", C.c_int32),
                ("stoch_k", C.c_int32), ("stoch_sk", C.c_int32), ("stoch_sd", C.c_int32),
                ("willr", C.c_int32), ("midprice", C.c_int32)]

    @classmethod
    def default(cls):
        return cls(30, 30, 30, 30, 20, 2.0, 2.0, 12, 26, 9, 14, 14, 14, 9, 3, 3, 14, 14)


def _col(x, ok=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if ok is not None:
        ok = np.ascontiguousarray(ok, dtype=np.uint8)
        assert ok.shape == x.shape
    return x, ok


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _outs(n, k):
    return [(np.empty(n, np.float64), np.empty(n, np.uint8)) for _ in range(k)]


def _ret(rc, outs):
    if rc != 0:
        raise OracleError(rc)
    res = [(v, o.astype(bool)) for v, o in outs]
    return res[0] if len(res) == 1 else tuple(res)


def _call(name, cols, ints, n_out, floats=()):
    """cols: list of (x, ok); ints: trailing int64 params; floats: trailing doubles after ints."""
    f = getattr(lib(), name)
    n = len(cols[0][0])
    args = []
    for x, ok in cols:
        args += [_p(x), _p(ok)]
    args.append(C.c_int64(n))
    args += [C.c_int64(int(i)) for i in ints]
    args += [C.c_double(float(d)) for d in floats]
    outs = _outs(n, n_out)
    for v, o in outs:
        args += [_p(v), _p(o)]
    f.restype = C.c_int
    return _ret(f(*args), outs)


def sma(x, p, ok=None): return _call("pqo_sma", [_col(x, ok)], [p], 1)
def ema(x, p, ok=None): return _call("pqo_ema", [_col(x, ok)], [p], 1)
def tema(x, p, ok=None): return _call("pqo_tema", [_col(x, ok)], [p], 1)
def trima(x, p, ok=None): return _call("pqo_trima", [_col(x, ok)], [p], 1)
def wma(x, p, ok=None): return _call("pqo_wma", [_col(x, ok)], [p], 1)
def dema(x, p, ok=None): return _call("pqo_dema", [_col(x, ok)], [p], 1)
def kama(x, p, ok=None): return _call("pqo_kama", [_col(x, ok)], [p], 1)
def t3(x, p, vfactor=0.0, ok=None): return _call("pqo_t3", [_col(x, ok)], [p], 1, floats=[vfactor])
def ma(x, p, matype=0, ok=None): return _call("pqo_ma", [_col(x, ok)], [p, matype], 1)
def midpoint(x, p=14, ok=None): return _call("pqo_midpoint", [_col(x, ok)], [p], 1)
def rsi(x, p=14, ok=None): return _call("pqo_rsi", [_col(x, ok)], [p], 1)
def mom(x, p=10, ok=None): return _call("pqo_mom", [_col(x, ok)], [p], 1)
def cmo(x, p=14, ok=None): return _call("pqo_cmo", [_col(x, ok)], [p], 1)


def rma(x, p):
    x, _ = _col(x)
    n = len(x)
    outs = _outs(n, 1)
    f = lib().pqo_rma
    f.restype = C.c_int
    return _ret(f(_p(x), C.c_int64(n), C.c_int64(p), _p(outs[0][0]), _p(outs[0][1])), outs)


def roc(x, p=10, kind=0, ok=None):
    x, ok = _col(x, ok)
    n = len(x)
    outs = _outs(n, 1)
    f = lib().pqo_roc
    f.restype = C.c_int
    return _ret(f(_p(x), _p(ok), C.c_int64(n), C.c_int64(p), C.c_int(kind), _p(outs[0][0]), _p(outs[0][1])), outs)


def bbands(x, p=20, up=2.0, dn=2.0, ok=None):
    return _call("pqo_bbands", [_col(x, ok)], [p], 3, floats=[up, dn])


def macd(x, fast=12, slow=26, signal=9, ok=None):
    return _call("pqo_macd", [_col(x, ok)], [fast, slow, signal], 3)


def midprice(h, l, p=14, hok=None, lok=None):
    return _call("pqo_midprice", [_col(h, hok), _col(l, lok)], [p], 1)


def trange(h, l, c, hok=None, lok=None, cok=None):
    return _call("pqo_trange", [_col(h, hok), _col(l, lok), _col(c, cok)], [], 1)


def atr(h, l, c, p=14, hok=None, lok=None, cok=None):
    return _call("pqo_atr", [_col(h, hok), _col(l, lok), _col(c, cok)], [p], 1)


def natr(h, l, c, p=14, hok=None, lok=None, cok=None):
    return _call("pqo_natr", [_col(h, hok), _col(l, lok), _col(c, cok)], [p], 1)


def willr(h, l, c, p=14, hok=None, lok=None, cok=None):
    return _call("pqo_willr", [_col(h, hok), _col(l, lok), _col(c, cok)], [p], 1)


def cci(h, l, c, p=14, hok=None, lok=None, cok=None):
    return _call("pqo_cci", [_col(h, hok), _col(l, lok), _col(c, cok)], [p], 1)


def obv(c, v, cok=None, vok=None):
    return _call("pqo_obv", [_col(c, cok), _col(v, vok)], [], 1)


def ad(h, l, c, v, hok=None, lok=None, cok=None, vok=None):
    return _call("pqo_ad", [_col(h, hok), _col(l, lok), _col(c, cok), _col(v, vok)], [], 1)


def adosc(h, l, c, v, fast=3, slow=10, hok=None, lok=None, cok=None, vok=None):
    return _call("pqo_adosc", [_col(h, hok), _col(l, lok), _col(c, cok), _col(v, vok)], [fast, slow], 1)


def mfi(h, l, c, v, p=14):
    return _call("pqo_mfi", [_col(h), _col(l), _col(c), _col(v)], [p], 1)


def stoch(h, l, c, fastk=5, slowk=3, slowk_matype=0, slowd=3, slowd_matype=0, hok=None, lok=None, cok=None):
    return _call("pqo_stoch", [_col(h, hok), _col(l, lok), _col(c, cok)],
                 [fastk, slowk, slowk_matype, slowd, slowd_matype], 2)


def stochf(h, l, c, fastk=5, fastd=3, fastd_matype=0, hok=None, lok=None, cok=None):
    return _call("pqo_stochf", [_col(h, hok), _col(l, lok), _col(c, cok)], [fastk, fastd, fastd_matype], 2)


def stochrsi(x, p=14, fastk=5, fastd=3, fastd_matype=0, ok=None):
    return _call("pqo_stochrsi", [_col(x, ok)], [p, fastk, fastd, fastd_matype], 2)


def macdext(x, fast=12, fastmatype=0, slow=26, slowmatype=0, signal=9, signalmatype=0, ok=None):
    return _call("pqo_macdext", [_col(x, ok)], [fast, fastmatype, slow, slowmatype, signal, signalmatype], 3)


def kdj(h, l, c, fastk=9, k=3, d=3):
    return _call("pqo_kdj", [_col(h), _col(l), _col(c)], [fastk, k, d], 3)


def dm(h, l, c, p=14):
    """-> dict plus_dm, minus_dm, dx (== the reference's plus_di), minus_di, adx, adxr: (values, ok) each."""
    res = _call("pqo_dm", [_col(h), _col(l), _col(c)], [p], 6)
    return dict(zip(("plus_dm", "minus_dm", "dx", "minus_di", "adx", "adxr"), res))


def trix(x, p=30, ok=None): return _call("pqo_trix", [_col(x, ok)], [p], 1)


def ultosc(h, l, c, p1=7, p2=14, p3=28):
    return _call("pqo_ultosc", [_col(h), _col(l), _col(c)], [p1, p2, p3], 1)


def aroon(h, l, p=14):
    """-> ((aroon_up, ok), (aroon_down, ok))"""
    return _call("pqo_aroon", [_col(h), _col(l)], [p], 2)


def donchian(h, l, p=20):
    h, _ = _col(h)
    l, _ = _col(l)
    n = len(h)
    outs = _outs(n, 2)
    f = lib().pqo_donchian
    f.restype = C.c_int
    rc = f(_p(h), _p(l), C.c_int64(n), C.c_int64(p), _p(outs[0][0]), _p(outs[0][1]), _p(outs[1][0]), _p(outs[1][1]))
    return _ret(rc, outs)


N_OUT = 21
OUTPUT_NAMES = ["sma", "ema", "tema", "trima", "bb_upper", "bb_middle", "bb_lower", "macd",
                "macd_signal", "macd_hist", "rsi", "trange", "atr", "natr", "obv", "ad",
                "kdj_k", "kdj_d", "kdj_j", "willr", "midprice"]


def suite_panel(c, h, l, v, params: SuiteParams | None = None, threads: int = 0):
    """c/h/l/v: float64 [n_symbols, pitch] C-contiguous (n_bars = pitch).  Returns
    (out [21, n_symbols, n_bars] float64, ok [21, n_symbols, n_bars] bool, threads_used)."""
    params = params or SuiteParams.default()
    c, h, l, v = (np.ascontiguousarray(a, dtype=np.float64) for a in (c, h, l, v))
    S, N = c.shape
    out = np.empty((N_OUT, S, N), np.float64)
    ok = np.empty((N_OUT, S, N), np.uint8)
    f = lib().pqo_suite_panel
    f.restype = C.c_int
    rc = f(_p(c), _p(h), _p(l), _p(v), C.c_int64(S), C.c_int64(N), C.c_int64(N), C.byref(params),
           _p(out), _p(ok), C.c_int(threads))
    if rc < 0:
        raise OracleError(rc)
    return out, ok.astype(bool), rc


# ---- candles (oracle/pq_candles.c): patterns, price transforms, bop ---------------------------------------------
_clib = None


def candles_lib():
    global _clib
    if _clib is None:
        build()
        _clib = C.CDLL(str(_SO_CANDLES))
    return _clib


def cdl(pattern: int, o, h, l, c, penetration: float = 0.3):
    """One candlestick pattern (id in the reference's order of definition) on one symbol -> int32 array."""
    o, h, l, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (o, h, l, c))
    out = np.empty(len(o), dtype=np.int32)
    rc = candles_lib().pqc_pattern(C.c_int(pattern), _p(o), _p(h), _p(l), _p(c), C.c_int64(len(o)), C.c_double(penetration), _p(out))
    if rc != 0:
        raise OracleError(rc)
    return out


def price(which: int, o, h, l, c):
    """avgprice 0, medprice 1, typprice 2, wclprice 3, bop 4 on null-free columns."""
    o, h, l, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (o, h, l, c))
    out = np.empty(len(o), dtype=np.float64)
    rc = candles_lib().pqc_price(C.c_int(which), _p(o), _p(h), _p(l), _p(c), C.c_int64(len(o)), _p(out))
    if rc != 0:
        raise OracleError(rc)
    return out


def candles_panel(o, h, l, c, penetration: float = 0.3, threads: int = 0):
    """All 61 patterns + 5 price outputs over [n_symbols, n_bars] arrays -> (int32 [61, S, N], float64 [5, S, N], threads)."""
    o, h, l, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (o, h, l, c))
    S, N = o.shape
    pat = np.empty((61, S, N), dtype=np.int32)
    pr = np.empty((5, S, N), dtype=np.float64)
    used = candles_lib().pqc_panel(_p(o), _p(h), _p(l), _p(c), C.c_int64(S), C.c_int64(N), C.c_int64(N), C.c_double(penetration),
                                   _p(pat), _p(pr), C.c_int(threads))
    if used <= 0:
        raise OracleError(used)
    return pat, pr, used


INFO_NAMES = ["price", "high", "low", "volume", "return_1d", "return_5d", "return_20d", "volatility", "ma_5", "ma_10",
              "ma_20", "volume_ratio", "amplitude"]


def info(c, h, l, v, start: int = 0):
    """Last-row reductions of one symbol (pqo_info) -> (float64[13], bool[13])."""
    c, h, l, v = (np.ascontiguousarray(x, dtype=np.float64) for x in (c, h, l, v))
    out = np.empty(13, dtype=np.float64)
    ok = np.zeros(13, dtype=np.uint8)
    rc = lib().pqo_info(_p(c), _p(h), _p(l), _p(v), C.c_int64(len(c)), C.c_int64(start), _p(out), _p(ok))
    if rc != 0:
        raise OracleError(rc)
    return out, ok.astype(bool)
