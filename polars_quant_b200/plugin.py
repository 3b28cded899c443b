"""Calls libpqb200.so's polars expression-plugin symbols (`_polars_plugin_<name>`, include/
pqb200_polars_plugin.h) exactly the way polars' plugin loader does -- Series exported through the Arrow C
Data Interface as polars-ffi SeriesExport structs, parameters as trailing length-1 literal Series and / or
pickled kwargs -- from pyarrow, so the reference-facing boundary can be exercised (and used) on a machine
without polars.  With polars installed, point the reference's shims at `LIB_PATH` instead:
`register_plugin_function(plugin_path=LIB_PATH, function_name="ema", args=[...], is_elementwise=False)`.
"""
from __future__ import annotations

import ctypes as C
import pickle

import numpy as np
import pyarrow as pa

from . import _native as N

LIB_PATH = N.LIB_PATH


class ArrowSchema(C.Structure):
    pass


class ArrowArray(C.Structure):
    pass


ArrowSchema._fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64),
                        ("n_children", C.c_int64), ("children", C.POINTER(C.POINTER(ArrowSchema))),
                        ("dictionary", C.POINTER(ArrowSchema)), ("release", C.c_void_p), ("private_data", C.c_void_p)]
ArrowArray._fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64),
                       ("n_children", C.c_int64), ("buffers", C.POINTER(C.c_void_p)),
                       ("children", C.POINTER(C.POINTER(ArrowArray))), ("dictionary", C.POINTER(ArrowArray)),
                       ("release", C.c_void_p), ("private_data", C.c_void_p)]


class SeriesExport(C.Structure):
    pass


SERIES_RELEASE = C.CFUNCTYPE(None, C.POINTER(SeriesExport))
SeriesExport._fields_ = [("field", C.POINTER(ArrowSchema)), ("arrays", C.POINTER(C.POINTER(ArrowArray))),
                         ("len", C.c_size_t), ("release", SERIES_RELEASE), ("private_data", C.c_void_p)]

_SCHEMA_RELEASE = C.CFUNCTYPE(None, C.POINTER(ArrowSchema))


class PluginError(RuntimeError):
    """A plugin call left `return_value` unset; the message is `_polars_plugin_get_last_error_message()`."""


def version() -> tuple[int, int]:
    L = N.lib()
    L._polars_plugin_get_version.restype = C.c_uint32
    v = L._polars_plugin_get_version()
    return v >> 16, v & 0xFFFF


def last_error() -> str:
    L = N.lib()
    L._polars_plugin_get_last_error_message.restype = C.c_char_p
    return (L._polars_plugin_get_last_error_message() or b"").decode("utf-8", "replace")


def _as_chunks(x, name):
    """-> (name, [pa.Array, ...]) for a data column or a literal parameter."""
    if isinstance(x, pa.ChunkedArray):
        return name, list(x.chunks) if x.num_chunks else [pa.array([], type=x.type)]
    if isinstance(x, pa.Array):
        return name, [x]
    if isinstance(x, np.ndarray):
        return name, [pa.array(x)]
    if isinstance(x, bool):
        raise TypeError("bool is not a numeric parameter")
    if isinstance(x, int):                       # polars materialises a Python int literal as Int32
        return "literal", [pa.array([x], type=pa.int32() if -2**31 <= x < 2**31 else pa.int64())]
    if isinstance(x, float):
        return "literal", [pa.array([x], type=pa.float64())]
    if x is None:
        return "literal", [pa.array([None], type=pa.int32())]
    return name, [pa.array(x)]


class _Exported:
    """One input Series as a SeriesExport whose release callback records that the callee called it."""

    def __init__(self, name, chunks):
        self.schema = ArrowSchema()
        pa.field(name, chunks[0].type)._export_to_c(C.addressof(self.schema))
        self.arrays = [ArrowArray() for _ in chunks]
        for a, c in zip(self.arrays, chunks):
            c._export_to_c(C.addressof(a))
        self.ptrs = (C.POINTER(ArrowArray) * len(chunks))(*[C.pointer(a) for a in self.arrays])
        self.released = False

        def _rel(p):
            self.released = True
            p.contents.release = SERIES_RELEASE()

        self._cb = SERIES_RELEASE(_rel)

    def fill(self, se: SeriesExport):
        se.field = C.pointer(self.schema)
        se.arrays = C.cast(self.ptrs, C.POINTER(C.POINTER(ArrowArray)))
        se.len = len(self.arrays)
        se.release = self._cb
        se.private_data = None

    def finish(self):
        """What polars' own release callback does afterwards: drop the field."""
        if self.schema.release:
            _SCHEMA_RELEASE(self.schema.release)(C.pointer(self.schema))
        return self.released and all(not a.release for a in self.arrays)


def call(function_name: str, args, kwargs: dict | None = None, names=None, check_consumed: bool = True):
    """One plugin call.  `args`: data columns (pyarrow arrays / chunked arrays / numpy) followed by optional
    literal parameters (Python int / float / None), in the order of the reference's Python shim; `kwargs`:
    pickled like polars does.  Returns a pyarrow Float64Array or StructArray; raises PluginError with the
    library's message when the call fails."""
    L = N.lib()
    fn = getattr(L, "_polars_plugin_" + function_name)
    fn.restype = None
    fn.argtypes = [C.POINTER(SeriesExport), C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(SeriesExport), C.c_void_p]
    exported = []
    for i, x in enumerate(args):
        nm, chunks = _as_chunks(x, (names[i] if names and i < len(names) else "column_%d" % i))
        exported.append(_Exported(nm, chunks))
    inputs = (SeriesExport * max(len(exported), 1))()
    for se, ex in zip(inputs, exported):
        ex.fill(se)
    kw = pickle.dumps(kwargs, protocol=5) if kwargs else b""
    ret = SeriesExport()
    fn(inputs, len(exported), kw if kw else None, len(kw), C.byref(ret), None)
    consumed = all([ex.finish() for ex in exported])
    if not ret.private_data:
        raise PluginError(last_error())
    if check_consumed and not consumed:
        raise AssertionError("the plugin did not release its inputs")
    assert ret.len == 1
    out = pa.Array._import_from_c(C.addressof(ret.arrays[0].contents), C.addressof(ret.field.contents))
    field_name = None
    ret.release(C.byref(ret))
    return out


def output_field(function_name: str, input_fields) -> pa.Field:
    """`_polars_plugin_field_<name>`: the output field polars asks for at planning time."""
    L = N.lib()
    fn = getattr(L, "_polars_plugin_field_" + function_name)
    fn.restype = None
    fn.argtypes = [C.POINTER(ArrowSchema), C.c_size_t, C.POINTER(ArrowSchema)]
    ins = (ArrowSchema * max(len(input_fields), 1))()
    for s, f in zip(ins, input_fields):
        f._export_to_c(C.addressof(s))
    out = ArrowSchema()
    fn(ins, len(input_fields), C.byref(out))
    for s in ins[:len(input_fields)]:
        if s.release:
            _SCHEMA_RELEASE(s.release)(C.pointer(s))
    if not out.release:
        raise PluginError(last_error())
    return pa.Field._import_from_c(C.addressof(out))
