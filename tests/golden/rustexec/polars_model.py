"""Models of the polars / std containers the reference's src/talib/*.rs touches.  TEST INFRASTRUCTURE.

Only behaviour the reference's text relies on is modelled, each item with the upstream semantics it
restates (polars 0.53 / polars-arrow, Rust std):

* `Option` is `Some(v)` / `None` (Python None); `Result` is `Ok(v)` / `Err(msg)`.
* `Float64Chunked`: a list of chunks (`PrimArray`: values + optional validity); arithmetic between
  chunked arrays / scalars is element-wise IEEE f64 with null propagation (polars' arithmetic kernels);
  `shift(k)` fills with nulls; `cont_slice()` is `Ok` only for ONE chunk WITHOUT nulls
  (polars-core `ChunkedArray::cont_slice`: "chunked array is not contiguous" otherwise).
* `PrimitiveChunkedBuilder::finish()` yields one chunk whose validity is present only if a null was
  appended (MutablePrimitiveArray -> PrimitiveArray drops an all-set bitmap).
* `VecDeque`, `ArrayVec` (fixed length: out-of-range index panics), slices / `Vec` as Python lists.
* f64 methods: `max` / `min` ignore a NaN operand (IEEE maxNum/minNum, as Rust documents), `mul_add`
  is a true fused multiply-add (libm `fma`), `powi` multiplies by squaring (compiler-rt `__powidf2`).
* A Rust panic (index out of bounds, `unwrap()` on `None`, ...) raises `RustPanic`; the crate is built
  with `panic = "abort"` (Cargo.toml:21), so in the reference it kills the process.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.fma.restype = ctypes.c_double
_libm.fma.argtypes = [ctypes.c_double] * 3

F64_MAX = 1.7976931348623157e308
U64 = 1 << 64


class RustPanic(Exception):
    pass


class Some:
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v

    def __repr__(self):
        return f"Some({self.v!r})"

    def __eq__(self, other):
        return isinstance(other, Some) and other.v == self.v

    def __hash__(self):
        return hash(("Some", self.v))


class Ok:
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v

    def __repr__(self):
        return f"Ok({self.v!r})"


class Err:
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v

    def __repr__(self):
        return f"Err({self.v!r})"


# ---- f64 / integer primitives ---------------------------------------------------------------------
def fma(a, b, c):
    return _libm.fma(a, b, c)


def fdiv(a, b):
    """IEEE 754 division (Python raises on a zero divisor)."""
    if b == 0.0:
        if a != a or a == 0.0:
            return math.nan
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf
    return a / b


def fsqrt(a):
    if a != a:
        return a
    if a < 0.0:
        return math.nan
    return math.sqrt(a)


def fmax(a, b):
    """f64::max: "if one of the arguments is NaN, then the other argument is returned"."""
    if a != a:
        return b
    if b != b:
        return a
    return a if a >= b else b


def fmin(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a <= b else b


def powi(a, n):
    """compiler-rt __powidf2: square-and-multiply; 1/x for negative exponents."""
    recip = n < 0
    n = abs(n)
    r = 1.0
    while True:
        if n & 1:
            r *= a
        n //= 2
        if n == 0:
            break
        a *= a
    return fdiv(1.0, r) if recip else r


def wrap_usize(v):
    """Release-profile integer arithmetic wraps (overflow checks are off with `cargo build --release`)."""
    return v % U64


# ---- std containers -----------------------------------------------------------------------------------
class VecDeque:
    def __init__(self):
        self.d = []

    def push_back(self, v):
        self.d.append(v)

    def pop_front(self):
        return Some(self.d.pop(0)) if self.d else None

    def pop_back(self):
        return Some(self.d.pop()) if self.d else None

    def front(self):
        return Some(self.d[0]) if self.d else None

    def back(self):
        return Some(self.d[-1]) if self.d else None

    def len(self):
        return len(self.d)

    def is_empty(self):
        return not self.d


class ArrayVec(list):
    """arrayvec::ArrayVec built with `from([x; N])`: length N, indexing past it panics."""


class RustIter:
    """A lazy iterator adaptor chain over a Python iterable."""

    def __init__(self, it):
        self.it = iter(it)

    def __iter__(self):
        return self.it


# ---- polars containers --------------------------------------------------------------------------------
class Bitmap:
    def __init__(self, bits):
        self.bits = bits

    def get_bit(self, i):
        if i >= len(self.bits):
            raise RustPanic("Bitmap::get_bit out of bounds")
        return self.bits[i]


class Buffer:
    def __init__(self, values):
        self.v = values

    def as_slice(self):
        return self.v


class PrimArray:
    """One Arrow chunk: `values` (list of float, slots under nulls hold arbitrary data) and `validity`
    (list of bool) or None."""

    def __init__(self, values, validity=None):
        self.vals = values
        self.valid = validity

    def len(self):
        return len(self.vals)

    def values(self):
        return Buffer(self.vals)

    def validity(self):
        return None if self.valid is None else Some(Bitmap(self.valid))

    def is_null(self, i):
        return self.valid is not None and not self.valid[i]

    def value(self, i):
        if i >= len(self.vals):
            raise RustPanic("PrimitiveArray::value out of bounds")
        return self.vals[i]

    def null_count(self):
        return 0 if self.valid is None else self.valid.count(False)

    def opt_items(self):
        if self.valid is None:
            return [Some(v) for v in self.vals]
        return [Some(v) if ok else None for v, ok in zip(self.vals, self.valid)]


class ChunkedArray:
    """Float64Chunked / Int64Chunked / Int32Chunked."""

    def __init__(self, name, chunks, dtype="Float64"):
        self.name, self.chunks, self.dtype = name, chunks, dtype

    # construction
    @staticmethod
    def from_options(name, items, dtype="Float64"):
        vals = [0.0 if it is None else it.v for it in items]
        valid = [it is not None for it in items]
        return ChunkedArray(name, [PrimArray(vals, None if all(valid) else valid)], dtype)

    @staticmethod
    def from_values(name, vals, dtype="Float64"):
        return ChunkedArray(name, [PrimArray(list(vals), None)], dtype)

    @staticmethod
    def full_null(name, n, dtype="Float64"):
        return ChunkedArray(name, [PrimArray([0.0] * n, [False] * n)], dtype)

    def opt_items(self):
        if getattr(self, "_items", None) is None:        # chunks are immutable once built
            out = []
            for c in self.chunks:
                out.extend(c.opt_items())
            self._items = out
        return self._items

    # the methods the reference calls
    def len(self):
        return sum(c.len() for c in self.chunks)

    def downcast_iter(self):
        return RustIter(self.chunks)

    def clone(self):
        return ChunkedArray(self.name, self.chunks, self.dtype)

    def with_name(self, name):
        return ChunkedArray(name, self.chunks, self.dtype)

    def rename(self, name):
        self.name = name

    def into_series(self):
        return Series(self)

    def rechunk(self):
        return ChunkedArray.from_options(self.name, self.opt_items(), self.dtype)

    def cont_slice(self):
        if len(self.chunks) == 1 and self.chunks[0].null_count() == 0:
            return Ok(self.chunks[0].vals)
        return Err("ComputeError: chunked array is not contiguous")

    def get(self, i):
        items = self.opt_items()
        return items[i] if i < len(items) else None

    def shift(self, k):
        items = self.opt_items()
        n = len(items)
        if k >= 0:
            items = [None] * min(k, n) + items[:max(n - k, 0)]
        else:
            items = items[min(-k, n):] + [None] * min(-k, n)
        return ChunkedArray.from_options(self.name, items, self.dtype)

    def into_iter(self):
        return RustIter(self.opt_items())

    iter = into_iter

    # D2: momentum.rs indexes the result of calc_ema / calc_sma like a Vec<Option<f64>>
    def index(self, i):
        items = self.opt_items()
        if i >= len(items):
            raise RustPanic("index out of bounds")
        return items[i]


def chunked_binop(op, a, b):
    """polars arithmetic kernels: element-wise, null if either side is null; a numeric scalar is
    converted to f64 (NumCast)."""
    fn = {"+": lambda x, y: x + y, "-": lambda x, y: x - y, "*": lambda x, y: x * y, "/": fdiv}[op]
    if isinstance(a, ChunkedArray) and isinstance(b, ChunkedArray):
        ia, ib = a.opt_items(), b.opt_items()
        if len(ia) != len(ib):
            if len(ib) == 1:
                ib = ib * len(ia)
            elif len(ia) == 1:
                ia = ia * len(ib)
            else:
                raise RustPanic("ShapeMismatch in chunked arithmetic")
        out = [Some(fn(x.v, y.v)) if x is not None and y is not None else None for x, y in zip(ia, ib)]
        return ChunkedArray.from_options(a.name, out)
    if isinstance(a, ChunkedArray):
        s = float(b)
        return ChunkedArray.from_options(a.name, [Some(fn(x.v, s)) if x is not None else None for x in a.opt_items()])
    s = float(a)
    return ChunkedArray.from_options(b.name, [Some(fn(s, y.v)) if y is not None else None for y in b.opt_items()])


class Builder:
    """PrimitiveChunkedBuilder::<Float64Type>."""

    def __init__(self, name, _capacity=0):
        self.name, self.vals, self.valid = name, [], []

    def append_value(self, v):
        self.vals.append(v)
        self.valid.append(True)

    def append_null(self):
        self.vals.append(0.0)
        self.valid.append(False)

    def append_option(self, o):
        if o is None:
            self.append_null()
        else:
            self.append_value(o.v)

    def finish(self):
        return ChunkedArray(self.name, [PrimArray(self.vals, None if all(self.valid) else self.valid)])


class StructChunked:
    def __init__(self, name, fields):
        self.name, self.fields = name, fields

    def into_series(self):
        return Series(self)


class Series:
    def __init__(self, inner):
        self.inner = inner

    @property
    def dtype(self):
        return "Struct" if isinstance(self.inner, StructChunked) else self.inner.dtype

    def name(self):
        return self.inner.name

    def len(self):
        return self.inner.len()

    def cast(self, dtype):
        ca = self.inner
        if dtype == ca.dtype:
            return Ok(Series(ca))
        if dtype == "Float64":
            return Ok(Series(ChunkedArray(ca.name, [PrimArray([float(v) for v in c.vals], c.valid) for c in ca.chunks])))
        if dtype in ("Int64", "Int32"):
            return Ok(Series(ChunkedArray(ca.name, [PrimArray([int(v) for v in c.vals], c.valid) for c in ca.chunks], dtype)))
        return Err(f"InvalidOperation: cast to {dtype}")

    def rechunk(self):
        return Series(self.inner.rechunk())

    def f64(self):
        if self.dtype != "Float64":
            return Err(f"SchemaMismatch: invalid series dtype: expected `Float64`, got `{self.dtype}`")
        return Ok(self.inner)

    def i64(self):
        if self.dtype != "Int64":
            return Err(f"SchemaMismatch: invalid series dtype: expected `Int64`, got `{self.dtype}`")
        return Ok(self.inner)

    def i32(self):
        if self.dtype != "Int32":
            return Err(f"SchemaMismatch: invalid series dtype: expected `Int32`, got `{self.dtype}`")
        return Ok(self.inner)

    def into_series(self):
        return self

    def clone(self):
        return self
