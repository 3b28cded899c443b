"""Candle engine binding (include/pqb200.h "candle engine"): the reference's 61 cdl* patterns, the four
price transforms and BOP from one fused pass over an open / high / low / close panel."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .panel import Engine, get_engine


def pattern_names() -> list[str]:
    return [N.lib().pqb_pattern_name(k).decode() for k in range(N.N_PATTERNS)]


def default_params(patterns: int | None = None, prices: int | None = None, penetration: float | None = None) -> N.CandleParams:
    p = N.CandleParams()
    N.lib().pqb_candle_params_default(C.byref(p))
    if patterns is not None:
        p.patterns = patterns
    if prices is not None:
        p.prices = prices
    if penetration is not None:
        for f in ("pen_darkcloudcover", "pen_eveningdojistar", "pen_eveningstar", "pen_morningdojistar",
                  "pen_morningstar", "pen_piercing"):
            setattr(p, f, float(penetration))
    return p


class CandlePanel:
    """open / high / low / close of `n_symbols` x `n_bars` on one B200 + the planes of the masked outputs."""

    def __init__(self, n_symbols: int, n_bars: int, engine: Engine | None = None, patterns_mask: int | None = None,
                 prices_mask: int | None = None, host_staging: bool = True):
        self.engine = engine or get_engine(0)
        self.n_symbols, self.n_bars = int(n_symbols), int(n_bars)
        self.patterns_mask = (1 << N.N_PATTERNS) - 1 if patterns_mask is None else int(patterns_mask)
        self.prices_mask = (1 << N.N_PRICES) - 1 if prices_mask is None else int(prices_mask)
        self._h = C.c_void_p()
        N.check(N.lib().pqb_candles_create(self.engine._h, self.n_symbols, self.n_bars, self.patterns_mask,
                                           self.prices_mask, 1 if host_staging else 0, C.byref(self._h)))
        self.pitch = N.lib().pqb_candles_pitch(self._h)
        self.words_per_row = (self.n_bars + 31) // 32

    def close(self):
        if self._h:
            N.lib().pqb_candles_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, ptr, dtype, cols):
        if not ptr:
            raise ValueError("this plane was not allocated")
        buf = (C.c_char * (self.n_symbols * cols * np.dtype(dtype).itemsize)).from_address(ptr)
        buf._owner = self                  # the view keeps the panel (and its pinned planes) alive, like Panel._view
        return np.frombuffer(buf, dtype=dtype).reshape(self.n_symbols, cols)

    def host_field(self, f: int) -> np.ndarray:
        return self._view(N.lib().pqb_candles_host_field(self._h, f), np.float64, self.pitch)[:, :self.n_bars]

    def set_fields(self, open, high, low, close, validity=None):
        """[n_symbols, n_bars] arrays; `validity`: optional dict field-index -> bool array (False = null)."""
        for f, a in enumerate((open, high, low, close)):
            a = np.ascontiguousarray(a, dtype=np.float64)
            assert a.shape == (self.n_symbols, self.n_bars)
            ok = None if validity is None else validity.get(f)
            if ok is None:
                self.host_field(f)[:] = a
            else:
                for s in range(self.n_symbols):
                    bits = np.packbits(np.asarray(ok[s], dtype=bool), bitorder="little")
                    N.check(N.lib().pqb_candles_set_column(self._h, s, f, a[s].ctypes.data_as(C.c_void_p),
                                                           bits.ctypes.data_as(C.c_void_p), 0, self.n_bars))

    def run_host(self, params: N.CandleParams | None = None):
        p = params or default_params(self.patterns_mask, self.prices_mask)
        N.check(N.lib().pqb_candles_run_host(self._h, C.byref(p)))

    def run(self, params: N.CandleParams | None = None):
        p = params or default_params(self.patterns_mask, self.prices_mask)
        N.check(N.lib().pqb_candles_run(self._h, C.byref(p)))

    def pattern(self, k: int) -> np.ndarray:
        """Int32 [n_symbols, n_bars] view of the host plane of pattern k (after run_host)."""
        return self._view(N.lib().pqb_candles_host_pattern(self._h, k), np.int32, self.pitch)[:, :self.n_bars]

    def price(self, k: int):
        """(values [n_symbols, n_bars], bool validity) of price output k (after run_host)."""
        v = self._view(N.lib().pqb_candles_host_price(self._h, k), np.float64, self.pitch)[:, :self.n_bars]
        w = self._view(N.lib().pqb_candles_host_price_validity(self._h, k), np.uint8, self.words_per_row * 4)
        ok = np.unpackbits(w, axis=1, bitorder="little")[:, :self.n_bars].astype(bool)
        return v, ok

    def fill_synthetic(self, seed: int = 1, to_host: bool = False):
        N.check(N.lib().pqb_candles_fill_synthetic(self._h, seed, 1 if to_host else 0))

    def fill_random_walk(self, seed: int = 0xC0FFEE, sigma: float = 0.02, to_host: bool = False):
        N.check(N.lib().pqb_candles_fill_random_walk(self._h, seed, sigma, 1 if to_host else 0))

    def time_device(self, params=None, warmup: int = 3, iters: int = 10) -> float:
        """Milliseconds of `iters` back-to-back fused candle launches on the device-resident panel."""
        p = params or default_params(self.patterns_mask, self.prices_mask)
        ms = C.c_float()
        N.check(N.lib().pqb_candles_time(self._h, C.byref(p), warmup, iters, C.byref(ms)))
        return float(ms.value)


def _col(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, N.Col(a.ctypes.data_as(C.c_void_p), None, 0, a.size)


def cdl(pattern, open, high, low, close, penetration: float = 0.3, engine: Engine | None = None) -> np.ndarray:
    """One reference plugin call: pattern name or id on one symbol's columns -> Int32 array."""
    e = engine or get_engine(0)
    k = pattern if isinstance(pattern, int) else N.lib().pqb_pattern_index(pattern.encode())
    keep = [_col(x) for x in (open, high, low, close)]
    out = np.zeros(keep[0][0].size, dtype=np.int32)
    N.check(N.lib().pqb_cdl(e._h, k, *[C.byref(c) for _, c in keep], float(penetration), out.ctypes.data_as(C.c_void_p)))
    return out
