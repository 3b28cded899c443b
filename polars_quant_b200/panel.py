"""Panel-level host API: the new layer between the reference's Python shims (L4) and its
per-column kernels (L1) -- SURVEY.md section 1.  A Panel packs the `{symbol}_{field}` f64
columns of a wide DataFrame into the GPU layout [field][symbol][pitch], runs the fused
indicator suite in one launch and hands the 21 output columns back with Arrow validity."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


class Engine:
    """One per GPU (pqb_engine)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        N.check(N.lib().pqb_engine_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            N.lib().pqb_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def flush_l2(self):
        N.check(N.lib().pqb_flush_l2(self._h))

    def local_cpus(self):
        """CPUs local to this engine's GPU (NVML affinity): where the engine pins its staging and runs its intake threads."""
        buf = (C.c_int * 1024)()
        n = N.lib().pqb_engine_local_cpus(self._h, buf, 1024)
        return [buf[i] for i in range(min(n, 1024))]


_engines: dict[int, Engine] = {}


def get_engine(device: int = 0) -> Engine:
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]


class Panel:
    FIELDS = {"close": N.CLOSE, "high": N.HIGH, "low": N.LOW, "volume": N.VOLUME}

    def __init__(self, n_symbols: int, n_bars: int, engine: Engine | None = None, fields_mask: int = 0xF,
                 outputs_mask: int = (1 << N.N_SUITE_OUTPUTS) - 1, host_staging: bool = True):
        self.engine = engine or get_engine(0)
        self.n_symbols, self.n_bars = int(n_symbols), int(n_bars)
        self.outputs_mask = outputs_mask
        self.host_staging = host_staging
        self._h = C.c_void_p()
        N.check(N.lib().pqb_panel_create(self.engine._h, n_symbols, n_bars, fields_mask, outputs_mask,
                                         int(host_staging) if host_staging else 0, C.byref(self._h)))
        self.pitch = N.lib().pqb_panel_pitch(self._h)
        self.validity_pitch = N.lib().pqb_panel_validity_pitch(self._h)

    @classmethod
    def _borrowed(cls, handle, engine, n_symbols, n_bars, outputs_mask, host_staging, owner):
        """A Panel over a native panel owned by another object (LongPanel, WindowPanel): same methods, never destroys."""
        self = cls.__new__(cls)
        self.engine, self.n_symbols, self.n_bars = engine, int(n_symbols), int(n_bars)
        self.outputs_mask, self.host_staging = outputs_mask, host_staging
        self._h = C.c_void_p(handle)
        self._owner = owner
        self.pitch = N.lib().pqb_panel_pitch(self._h)
        self.validity_pitch = N.lib().pqb_panel_validity_pitch(self._h)
        return self

    def close(self):
        if self._h and getattr(self, "_owner", None) is None:
            N.lib().pqb_panel_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host staging views (zero-copy numpy over pinned memory) ----
    def _view(self, ptr, ctype, count, dtype):
        """numpy view of `count` elements of pinned staging.  The view (and every array sliced from it) holds a
        reference to this Panel through its base buffer, so the pinned planes outlive a dropped Panel object -- they
        return to the engine's pool only when the last view is gone.  An explicit close() still invalidates views."""
        buf = (ctype * count).from_address(ptr)
        buf._owner = self
        return np.frombuffer(buf, dtype=dtype)

    def host_field(self, field) -> np.ndarray:
        f = self.FIELDS[field] if isinstance(field, str) else field
        ptr = N.lib().pqb_panel_host_field(self._h, f)
        if not ptr:
            raise ValueError("panel has no host staging for field %r" % (field,))
        return self._view(ptr, C.c_double, self.n_symbols * self.pitch, np.float64).reshape(self.n_symbols, self.pitch)

    def host_output(self, k: int) -> np.ndarray:
        ptr = N.lib().pqb_panel_host_output(self._h, k)
        if not ptr:
            raise ValueError("output %d is not allocated (or the panel has no host staging)" % k)
        return self._view(ptr, C.c_double, self.n_symbols * self.pitch, np.float64).reshape(self.n_symbols, self.pitch)[:, :self.n_bars]

    def host_validity(self, k: int) -> np.ndarray:
        """bool [n_symbols, n_bars] unpacked from the Arrow LSB-first bitmaps."""
        ptr = N.lib().pqb_panel_host_validity(self._h, k)
        if not ptr:
            raise ValueError("output %d is not allocated (or the panel has no host staging)" % k)
        bits = self._view(ptr, C.c_uint8, self.n_symbols * self.validity_pitch, np.uint8).reshape(self.n_symbols, self.validity_pitch)
        return np.unpackbits(bits, axis=1, bitorder="little")[:, :self.n_bars].astype(bool)

    # ---- loading ----
    def set_fields(self, close=None, high=None, low=None, volume=None, starts=None):
        """Each argument: float64 [n_symbols, n_bars].  Written straight into pinned staging."""
        for name, a in (("close", close), ("high", high), ("low", low), ("volume", volume)):
            if a is None:
                continue
            a = np.asarray(a, dtype=np.float64)
            if a.shape != (self.n_symbols, self.n_bars):
                raise ValueError("%s has shape %r, panel is %r" % (name, a.shape, (self.n_symbols, self.n_bars)))
            v = self.host_field(name)
            v[:, :self.n_bars] = a
            v[:, self.n_bars:] = 0.0
            # (dense values written past the column intake: whatever validity earlier set_column calls recorded is void)
            N.check(N.lib().pqb_panel_clear_validity(self._h, self.FIELDS[name]))
        if starts is not None:
            s = np.ascontiguousarray(starts, dtype=np.int32)
            N.check(N.lib().pqb_panel_set_starts(self._h, s.ctypes.data_as(C.c_void_p)))

    def set_column(self, symbol: int, field, values: np.ndarray, validity: np.ndarray | None = None, offset: int = 0):
        """One Arrow column: values float64, validity = packed LSB-first bitmap (uint8) or None."""
        f = self.FIELDS[field] if isinstance(field, str) else field
        values = np.ascontiguousarray(values, dtype=np.float64)
        vp = None if validity is None else np.ascontiguousarray(validity, dtype=np.uint8).ctypes.data_as(C.c_void_p)
        N.check(N.lib().pqb_panel_set_column(self._h, symbol, f, values.ctypes.data_as(C.c_void_p), vp, offset,
                                             len(values) - offset))

    def fill_synthetic(self, seed: int = 0xC0FFEE, sigma: float = 0.02, to_host: bool = False):
        N.check(N.lib().pqb_panel_fill_synthetic(self._h, seed, sigma, 1 if to_host else 0))

    @staticmethod
    def col_refs(columns):
        """[(symbol, field, values float64 ndarray, validity packed-bitmap uint8 ndarray | None), ...] -> a ctypes array
        of pqb_col_ref (+ the list of arrays it borrows from: keep it alive as long as the refs are used)."""
        refs = (N.ColRef * len(columns))()
        keep = []
        for r, (symbol, field, values, validity) in zip(refs, columns):
            v = np.ascontiguousarray(values, dtype=np.float64)
            keep.append(v)
            r.values, r.offset, r.len, r.symbol = v.ctypes.data, 0, v.size, symbol
            r.field = Panel.FIELDS[field] if isinstance(field, str) else int(field)
            if validity is not None:
                b = np.ascontiguousarray(validity, dtype=np.uint8)
                keep.append(b)
                r.validity = b.ctypes.data
        return refs, keep

    @staticmethod
    def field_refs(close=None, high=None, low=None, volume=None):
        """pqb_col_ref array for dense [n_symbols, n_bars] float64 field matrices (C-contiguous rows), built without a
        Python loop: one ref per (symbol, field) pointing into the caller's own memory."""
        mats = [(f, np.ascontiguousarray(a, dtype=np.float64)) for f, a in
                ((N.CLOSE, close), (N.HIGH, high), (N.LOW, low), (N.VOLUME, volume)) if a is not None]
        S, nb = mats[0][1].shape
        rec = np.zeros(S * len(mats), dtype=np.dtype([("values", np.uint64), ("validity", np.uint64), ("offset", np.int64),
                                                      ("len", np.int64), ("symbol", np.int64), ("field", np.int32),
                                                      ("reserved", np.int32)]))
        for i, (f, a) in enumerate(mats):
            sl = rec[i * S:(i + 1) * S]
            sl["values"] = a.ctypes.data + np.arange(S, dtype=np.uint64) * np.uint64(a.strides[0])
            sl["len"], sl["symbol"], sl["field"] = nb, np.arange(S), f
        return rec, [a for _, a in mats]

    def set_columns(self, refs, n=None, threads: int = 0):
        """Batch intake (pqb_panel_set_columns): `refs` from col_refs() / field_refs()."""
        ptr, n = (refs.ctypes.data, len(refs)) if isinstance(refs, np.ndarray) else (C.addressof(refs), len(refs) if n is None else n)
        N.check(N.lib().pqb_panel_set_columns(self._h, ptr, n, threads))

    def run_columns(self, refs, params: N.SuiteParams | None = None, threads: int = 0):
        """Caller-owned columns in, results in the pinned result planes: intake pipelined with H2D / suite / D2H
        (pqb_suite_run_columns)."""
        params = params or N.default_params()
        ptr, n = (refs.ctypes.data, len(refs)) if isinstance(refs, np.ndarray) else (C.addressof(refs), len(refs))
        N.check(N.lib().pqb_suite_run_columns(self._h, C.byref(params), ptr, n, threads))

    def export_arrow(self, outputs_mask: int = 0, symbol_names=None):
        """Every result column of the last run as ONE pyarrow RecordBatch (`{symbol}_{output}` Float64 columns aliasing
        the pinned result planes; the batch keeps the native panel alive)."""
        import pyarrow as pa
        arr, sch = N.ArrowArray(), N.ArrowSchema()
        names = None
        if symbol_names is not None:
            enc = [s.encode() for s in symbol_names]
            names = (C.c_char_p * len(enc))(*enc)
        N.check(N.lib().pqb_panel_export_arrow(self._h, outputs_mask, names, C.byref(arr), C.byref(sch)))
        return pa.RecordBatch._import_from_c(C.addressof(arr), C.addressof(sch))

    # ---- running ----
    def upload(self):
        N.check(N.lib().pqb_panel_upload(self._h))

    def download(self):
        N.check(N.lib().pqb_panel_download(self._h))

    def sync(self):
        N.check(N.lib().pqb_panel_sync(self._h))

    def run(self, params: N.SuiteParams | None = None):
        """Device-resident fused suite (async)."""
        params = params or N.default_params()
        N.check(N.lib().pqb_suite_run(self._h, C.byref(params)))

    def run_host(self, params: N.SuiteParams | None = None, chunk_symbols: int = 0):
        """End-to-end: pinned staging -> device -> suite -> pinned staging (pipelined)."""
        params = params or N.default_params()
        N.check(N.lib().pqb_suite_run_host(self._h, C.byref(params), chunk_symbols))

    def compute(self, params: N.SuiteParams | None = None):
        """upload + run + download + sync; returns {name: (values, validity)}."""
        self.upload()
        self.run(params)
        self.download()
        self.sync()
        return self.outputs()

    def outputs(self):
        res = {}
        for k, name in enumerate(N.OUTPUT_NAMES):
            if self.outputs_mask >> k & 1:
                res[name] = (self.host_output(k), self.host_validity(k))
        return res

    SIGNALS = ("macd_cross", "kdj_cross", "rsi_cross")

    def signals(self, oversold: float = 30.0, overbought: float = 70.0):
        """Crossover signals (int8 +1 / -1 / 0) from the outputs of the last run: {name: [n_symbols, n_bars] array}."""
        N.check(N.lib().pqb_signals_run(self._h, oversold, overbought))
        res = {}
        for q, name in enumerate(self.SIGNALS):
            ptr = N.lib().pqb_panel_host_signal(self._h, q)
            if not ptr:
                raise ValueError("signals need a panel with host staging")
            res[name] = self._view(ptr, C.c_int8, self.n_symbols * self.pitch, np.int8).reshape(self.n_symbols, self.pitch)[:, :self.n_bars]
        return res

    def ma_cross(self, fast_period: int = 10, slow_period: int = 20, ma_type: str = "sma"):
        """MA golden / death cross (README.md:876-905 `Strategy.ma`): int8 [n_symbols, n_bars], +1 where MA(close, fast)
        crosses above MA(close, slow), -1 where it crosses below.  close must be on the device (upload / run_host)."""
        kinds = {"sma": 0, "ema": 1}
        if ma_type not in kinds:
            raise ValueError("ma_type %r: 'sma' and 'ema' are built" % (ma_type,))
        N.check(N.lib().pqb_ma_cross_run(self._h, kinds[ma_type], fast_period, slow_period))
        ptr = N.lib().pqb_panel_host_signal(self._h, 3)
        if not ptr:
            raise ValueError("signals need a panel with host staging")
        return self._view(ptr, C.c_int8, self.n_symbols * self.pitch, np.int8).reshape(self.n_symbols, self.pitch)[:, :self.n_bars]

    def info(self):
        """Last-row reductions per symbol (the README's `Selector.info()` columns that come from close / high / low /
        volume; include/pqb200.h `pqb_panel_info`): {name: (float64[n_symbols], bool[n_symbols] valid)}.  The inputs
        must be on the device (upload / run_host / fill_synthetic)."""
        out = np.empty((len(N.INFO_NAMES), self.n_symbols), dtype=np.float64)
        ok = np.empty((len(N.INFO_NAMES), self.n_symbols), dtype=np.uint8)
        N.check(N.lib().pqb_panel_info(self._h, out.ctypes.data, ok.ctypes.data))
        return {name: (out[k], ok[k].astype(bool)) for k, name in enumerate(N.INFO_NAMES)}

    def last_launches(self) -> int:
        return N.lib().pqb_panel_last_launches(self._h)

    def tiled_shape(self):
        """(symbol blocks, padded bars per block) of the tiled device planes."""
        nb, bp = C.c_int64(), C.c_int64()
        N.check(N.lib().pqb_panel_tiled_shape(self._h, C.byref(nb), C.byref(bp)))
        return nb.value, bp.value

    # ---- measurement ----
    def time_device(self, params=None, warmup: int = 3, iters: int = 20):
        params = params or N.default_params()
        tot, fused, nl = C.c_float(), C.c_float(), C.c_int()
        N.check(N.lib().pqb_suite_time(self._h, C.byref(params), warmup, iters, C.byref(tot), C.byref(fused), C.byref(nl)))
        return tot.value, fused.value, nl.value

    def time_host(self, params=None, chunk_symbols: int = 0, warmup: int = 1, iters: int = 3):
        params = params or N.default_params()
        tot = C.c_float()
        N.check(N.lib().pqb_suite_time_host(self._h, C.byref(params), chunk_symbols, warmup, iters, C.byref(tot)))
        return tot.value


class SplitPanel:
    """Long rows, few symbols (BASELINE config 3): every row runs as `chunks` independent virtual symbols, each with
    `warmup` bars of real history in front (include/pqb200.h "time-split panels").  Chunk 0 is bit-exact; later
    chunks agree with the serial computation to < 1e-12 relative when warmup >= required_warmup(params)."""

    def __init__(self, n_symbols: int, n_bars: int, chunks: int, warmup: int, engine: Engine | None = None,
                 fields_mask: int = 0xF, outputs_mask: int = (1 << N.N_SUITE_OUTPUTS) - 1, host_staging: bool = True):
        self.engine = engine or get_engine(0)
        self.n_symbols, self.n_bars = int(n_symbols), int(n_bars)
        self.outputs_mask = outputs_mask
        self._h = C.c_void_p()
        N.check(N.lib().pqb_split_create(self.engine._h, n_symbols, n_bars, chunks, warmup, fields_mask, outputs_mask,
                                         1 if host_staging else 0, C.byref(self._h)))
        a, b, c, d = (C.c_int64() for _ in range(4))
        N.check(N.lib().pqb_split_shape(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        self.chunk_bars, self.warmup, self.virtual_symbols, self.virtual_bars = a.value, b.value, c.value, d.value

    @staticmethod
    def required_warmup(params: N.SuiteParams) -> int:
        return int(N.lib().pqb_split_required_warmup(C.byref(params)))

    def close(self):
        if self._h:
            N.lib().pqb_split_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_column(self, symbol: int, field, values: np.ndarray):
        f = Panel.FIELDS[field] if isinstance(field, str) else int(field)
        v = np.ascontiguousarray(values, dtype=np.float64)
        N.check(N.lib().pqb_split_set_column(self._h, symbol, f, v.ctypes.data_as(C.c_void_p), None, 0, v.size))

    def set_fields(self, close=None, high=None, low=None, volume=None):
        for f, a in ((N.CLOSE, close), (N.HIGH, high), (N.LOW, low), (N.VOLUME, volume)):
            if a is not None:
                for s in range(self.n_symbols):
                    self.set_column(s, f, a[s])

    def run_host(self, params: N.SuiteParams | None = None):
        params = params or N.default_params()
        N.check(N.lib().pqb_split_run_host(self._h, C.byref(params)))

    def get_output(self, symbol: int, k: int):
        v = np.empty(self.n_bars, dtype=np.float64)
        bits = np.zeros((self.n_bars + 7) // 8, dtype=np.uint8)
        N.check(N.lib().pqb_split_get_output(self._h, symbol, k, v.ctypes.data_as(C.c_void_p), bits.ctypes.data_as(C.c_void_p),
                                             self.n_bars))
        return v, np.unpackbits(bits, bitorder="little")[:self.n_bars].astype(bool)

    def fill_synthetic(self, seed: int = 0xC0FFEE, sigma: float = 0.0005):
        N.check(N.lib().pqb_split_fill_synthetic(self._h, seed, sigma))

    def time_device(self, params=None, warmup: int = 2, iters: int = 5):
        """(total ms, fused-kernel ms, launches per step) of `iters` passes over the device-resident virtual panel."""
        params = params or N.default_params()
        tot, fused, nl = C.c_float(), C.c_float(), C.c_int()
        inner = N.lib().pqb_split_panel(self._h)
        N.check(N.lib().pqb_suite_time(inner, C.byref(params), warmup, iters, C.byref(tot), C.byref(fused), C.byref(nl)))
        return tot.value, fused.value, nl.value


class MultiPanel:
    """Multi-GPU driver (pqb_multi_*): one shard = engine + panel per listed device, symbols split in
    contiguous whole-block ranges, one host thread per shard, no collective."""

    def __init__(self, n_symbols: int, n_bars: int, devices, fields_mask: int = 0xF,
                 outputs_mask: int = (1 << N.N_SUITE_OUTPUTS) - 1):
        self.n_symbols, self.n_bars, self.outputs_mask = int(n_symbols), int(n_bars), outputs_mask
        devs = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p()
        N.check(N.lib().pqb_multi_create(devs, len(devices), C.c_int64(n_symbols), C.c_int64(n_bars),
                                         C.c_uint32(fields_mask), C.c_uint64(outputs_mask), C.byref(self._h)))

    def close(self):
        if self._h:
            N.lib().pqb_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shards(self):
        out = []
        for i in range(N.lib().pqb_multi_shard_count(self._h)):
            dev, lo, hi, ph = C.c_int(), C.c_int64(), C.c_int64(), C.c_void_p()
            N.check(N.lib().pqb_multi_shard(self._h, i, C.byref(dev), C.byref(lo), C.byref(hi), C.byref(ph)))
            out.append((dev.value, lo.value, hi.value))
        return out

    def set_column(self, symbol: int, field, values, validity=None, offset: int = 0):
        f = Panel.FIELDS[field] if isinstance(field, str) else field
        values = np.ascontiguousarray(values, dtype=np.float64)
        vp = None if validity is None else np.ascontiguousarray(validity, dtype=np.uint8).ctypes.data_as(C.c_void_p)
        N.check(N.lib().pqb_multi_set_column(self._h, C.c_int64(symbol), f, values.ctypes.data_as(C.c_void_p), vp,
                                             C.c_int64(offset), C.c_int64(len(values) - offset)))

    def run_host(self, params=None):
        params = params or N.default_params()
        N.check(N.lib().pqb_multi_run_host(self._h, C.byref(params)))

    def get_output(self, symbol: int, k: int):
        v = np.empty(self.n_bars)
        b = np.zeros((self.n_bars + 7) // 8, np.uint8)
        N.check(N.lib().pqb_multi_get_output(self._h, C.c_int64(symbol), k, v.ctypes.data_as(C.c_void_p),
                                             b.ctypes.data_as(C.c_void_p), C.c_int64(self.n_bars)))
        return v, np.unpackbits(b, bitorder="little")[:self.n_bars].astype(bool)
