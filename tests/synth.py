"""Synthetic OHLCV panels: geometric random-walk close, high/low around
open/close, integer-valued log-normal volume.  numpy Philox, one stream per panel seed."""
from __future__ import annotations

import numpy as np


def ohlcv(n_symbols: int, n_bars: int, seed: int = 0xC0FFEE, sigma: float = 0.02):
    """Returns dict of float64 [n_symbols, n_bars] arrays: open, high, low, close, volume."""
    rng = np.random.Generator(np.random.Philox(seed))
    eps = rng.normal(0.0, sigma, size=(n_symbols, n_bars))
    close = 100.0 * np.exp(np.cumsum(eps, axis=1))
    open_ = np.empty_like(close)
    open_[:, 0] = 100.0
    open_[:, 1:] = close[:, :-1]
    up = np.abs(rng.normal(0.0, sigma / 2, size=close.shape))
    dn = np.abs(rng.normal(0.0, sigma / 2, size=close.shape))
    high = np.maximum(open_, close) * (1.0 + up)
    low = np.minimum(open_, close) * (1.0 - dn)
    volume = np.round(rng.lognormal(13.0, 1.0, size=close.shape))
    return {"open": open_, "high": high, "low": low, "close": close, "volume": volume}


def to_opt(x, ok=None):
    """numpy column (+ bool validity) -> list[float|None] for oracle/ref_py.py."""
    if ok is None:
        return [float(v) for v in x]
    return [float(v) if k else None for v, k in zip(x, ok)]


def from_opt(col):
    """list[float|None] -> (values with NaN at nulls, bool validity)."""
    ok = np.array([v is not None for v in col], dtype=bool)
    vals = np.array([np.nan if v is None else v for v in col], dtype=np.float64)
    return vals, ok
