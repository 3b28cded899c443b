"""Tree-walking evaluator for the syntax trees of rs_parse.py.  TEST INFRASTRUCTURE (see rs_parse.py).

Semantics that matter for bit-exact results:
* f64 arithmetic = Python float arithmetic (IEEE binary64, round to nearest even), `/` through
  polars_model.fdiv (IEEE results for zero divisors), `mul_add` through libm's fma; nothing is
  re-associated: the tree is evaluated in Rust's precedence and left-to-right order.
* integers are Python ints with release-profile wrap-around to usize on `+ - *` (every integer on this
  path is a usize: counters, periods, indices), truncating `/`, `%`.
* `&x`, `*x`, `&mut x` are transparent (values are shared by reference where Rust borrows them).
* `?` returns early from the enclosing fn / closure on `Err` / `None`.
"""
from __future__ import annotations

from . import polars_model as M
from .polars_model import Err, Ok, RustPanic, Some


class _Return(Exception):
    def __init__(self, v):
        self.v = v


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class RustRuntimeError(Exception):
    pass


class Struct:
    """An instance of a `#[derive(Deserialize)] struct` (the kwargs of a plugin function)."""

    def __init__(self, name, fields):
        self.__dict__.update(fields)
        self._name = name


class Closure:
    def __init__(self, interp, params, body, env):
        self.interp, self.params, self.body, self.env = interp, params, body, env

    def __call__(self, *args):
        env = Env(self.env)
        if len(self.params) != len(args):
            raise RustRuntimeError("closure arity")
        for pat, a in zip(self.params, args):
            if not self.interp.bind(pat, a, env):
                raise RustPanic("refutable closure parameter did not match")
        try:
            return self.interp.ev(self.body, env)
        except _Return as r:
            return r.v


class Env:
    __slots__ = ("vars", "up")

    def __init__(self, up=None):
        self.vars = {}
        self.up = up

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e
            e = e.up
        return None


class Function:
    def __init__(self, interp, node, module):
        self.interp, self.node, self.module = interp, node, module
        self.name = node[1]

    def __call__(self, *args):
        return self.interp.call_fn(self, list(args))


class Module:
    def __init__(self, name):
        self.name = name
        self.fns = {}
        self.structs = {}


def _is_f64_chunked_param(ty):
    while ty and ty[0] == "ref":
        ty = ty[1]
    return bool(ty) and ty[0] == "path" and ty[1] == "Float64Chunked"


class Interp:
    """Holds the parsed modules; `call(module, fn, *args)` runs a function."""

    def __init__(self):
        self.modules = {}
        self.extern = {}        # functions the reference calls but does not define (D1: calc_rma)
        self.trace = None

    # ---- loading -------------------------------------------------------------------------------
    def load(self, name, parsed):
        mod = Module(name)
        for fname, node in parsed["fns"].items():
            mod.fns[fname] = Function(self, node, mod)
        mod.structs = parsed["structs"]
        self.modules[name] = mod
        return mod

    def resolve_fn(self, module, name):
        if name in module.fns:
            return module.fns[name]
        for mod in self.modules.values():
            if name in mod.fns:
                return mod.fns[name]
        if name in self.extern:
            return self.extern[name]
        return None

    def make_struct(self, module, name, **fields):
        names = self.modules[module].structs[name]
        vals = {f: None for f in names}
        for k, v in fields.items():
            if k not in vals:
                raise RustRuntimeError(f"{name} has no field {k}")
            vals[k] = Some(v)
        return Struct(name, vals)

    def call(self, module, fn, *args):
        f = self.modules[module].fns[fn]
        return f(*args)

    # ---- functions -----------------------------------------------------------------------------
    def call_fn(self, fn, args):
        node = fn.node
        params = node[2]
        if len(params) != len(args):
            raise RustRuntimeError(f"{fn.name}: expected {len(params)} arguments, got {len(args)}")
        env = Env(None)
        env.vars["__module__"] = fn.module
        # D2 adapter: momentum.rs calls overlap's calc_ema / calc_sma with a `&[f64]` /
        # `&Vec<f64>` and uses the result as a Vec<Option<f64>>.  The slice is presented to the callee's own
        # text as a null-free single-chunk Float64Chunked; the result keeps list-like access (index / iter /
        # len) through ChunkedArray.index.  No arithmetic is added or changed.
        for k, ((pat, ty), a) in enumerate(zip(params, args)):
            if _is_f64_chunked_param(ty) and isinstance(a, list):
                a = M.ChunkedArray.from_values("", a)
            if not self.bind(pat, a, env):
                raise RustPanic("refutable parameter pattern")
        try:
            return self.ev(node[4], env)
        except _Return as r:
            return r.v

    # ---- patterns ------------------------------------------------------------------------------
    def bind(self, pat, v, env):
        k = pat[0]
        if k == "pbind":
            env.vars[pat[1]] = v
            return True
        if k == "pwild":
            return True
        if k == "ptuple":
            if not isinstance(v, tuple) or len(v) != len(pat[1]):
                return False
            return all(self.bind(p, x, env) for p, x in zip(pat[1], v))
        if k == "pnone":
            return v is None
        if k == "pctor":
            cls = {"Some": Some, "Ok": Ok, "Err": Err}[pat[1]]
            if not isinstance(v, cls):
                return False
            return self.bind(pat[2], v.v, env)
        if k == "plit":
            return type(v) is type(pat[1]) and v == pat[1] or (isinstance(v, (int, float)) and not isinstance(v, bool)
                                                               and not isinstance(pat[1], bool) and v == pat[1])
        raise RustRuntimeError(f"pattern {k}")

    # ---- evaluation ----------------------------------------------------------------------------
    def ev(self, n, env):
        return getattr(self, "ev_" + n[0])(n, env)

    def ev_lit(self, n, env):
        return n[1]

    def ev_paren(self, n, env):
        return self.ev(n[1], env)

    def ev_var(self, n, env):
        e = env.lookup(n[1])
        if e is not None:
            return e.vars[n[1]]
        module = self.module_of(env)
        f = self.resolve_fn(module, n[1])
        if f is not None:
            return f
        if n[1] in CTORS:
            return CTORS[n[1]]
        if n[1] == "None":
            return None
        raise RustRuntimeError(f"line {n[2]}: unresolved name `{n[1]}` (undefined in the reference snapshot?)")

    def module_of(self, env):
        e = env
        while e.up is not None:
            e = e.up
        return e.vars.get("__module__")

    def ev_path(self, n, env):
        p = n[1]
        if p in PATHS:
            return PATHS[p]
        last = p.rsplit("::", 1)[-1]
        f = self.resolve_fn(self.module_of(env), last)
        if f is not None and p.startswith(("crate::", "super::", "self::")):
            return f
        raise RustRuntimeError(f"line {n[2]}: unknown path `{p}`")

    def ev_tuple(self, n, env):
        return tuple(self.ev(x, env) for x in n[1])

    def ev_array(self, n, env):
        return [self.ev(x, env) for x in n[1]]

    def ev_repeat(self, n, env):
        v = self.ev(n[1], env)
        return [v] * self.ev(n[2], env)

    def ev_block(self, n, env):
        env = Env(env)
        for s in n[1]:
            k = s[0]
            if k == "let":
                v = self.ev(s[2], env) if s[2] is not None else None
                if not self.bind(s[1], v, env):
                    raise RustPanic(f"line {s[3]}: refutable pattern in let")
            elif k == "expr":
                self.ev(s[1], env)
            elif k == "fnitem":
                env.vars[s[1][1]] = Function(self, s[1], self.module_of(env))
        if n[2] is not None:
            return self.ev(n[2], env)
        return ()

    def ev_if(self, n, env):
        if self.ev(n[1], env):
            return self.ev(n[2], env)
        if n[3] is not None:
            return self.ev(n[3], env)
        return ()

    def ev_iflet(self, n, env):
        v = self.ev(n[2], env)
        inner = Env(env)
        if self.bind(n[1], v, inner):
            return self.ev(n[3], inner)
        if n[4] is not None:
            return self.ev(n[4], env)
        return ()

    def ev_match(self, n, env):
        v = self.ev(n[1], env)
        for pats, guard, body in n[2]:
            for pat in pats:
                inner = Env(env)
                if self.bind(pat, v, inner) and (guard is None or self.ev(guard, inner)):
                    return self.ev(body, inner)
        raise RustPanic(f"line {n[3]}: no match arm")

    def iterate(self, v):
        if isinstance(v, (list, tuple, M.RustIter)):
            return v
        if isinstance(v, range):
            return v
        if isinstance(v, M.ChunkedArray):
            return v.opt_items()
        raise RustRuntimeError(f"cannot iterate {type(v).__name__}")

    def ev_for(self, n, env):
        for item in self.iterate(self.ev(n[2], env)):
            inner = Env(env)
            if not self.bind(n[1], item, inner):
                raise RustPanic("refutable for pattern")
            try:
                self.ev(n[3], inner)
            except _Continue:
                continue
            except _Break:
                break
        return ()

    def ev_while(self, n, env):
        while self.ev(n[1], env):
            try:
                self.ev(n[2], env)
            except _Continue:
                continue
            except _Break:
                break
        return ()

    def ev_whilelet(self, n, env):
        while True:
            inner = Env(env)
            if not self.bind(n[1], self.ev(n[2], env), inner):
                break
            try:
                self.ev(n[3], inner)
            except _Continue:
                continue
            except _Break:
                break
        return ()

    def ev_loop(self, n, env):
        while True:
            try:
                self.ev(n[1], env)
            except _Continue:
                continue
            except _Break:
                break
        return ()

    def ev_return(self, n, env):
        raise _Return(self.ev(n[1], env) if n[1] is not None else ())

    def ev_break(self, n, env):
        raise _Break()

    def ev_continue(self, n, env):
        raise _Continue()

    def ev_closure(self, n, env):
        return Closure(self, n[1], n[2], env)

    def ev_range(self, n, env):
        lo, hi = self.ev(n[1], env), self.ev(n[2], env)
        return range(lo, hi + 1 if n[3] else hi)

    def ev_neg(self, n, env):
        v = self.ev(n[1], env)
        if isinstance(v, float):
            return -v
        return -v               # negative integer literals (`-100`); usize negation does not occur

    def ev_not(self, n, env):
        v = self.ev(n[1], env)
        if isinstance(v, bool):
            return not v
        raise RustRuntimeError("`!` on a non-bool")

    def ev_cast(self, n, env):
        v = self.ev(n[1], env)
        ty = n[2]
        if ty in ("f64", "f32"):
            return float(v)
        if ty in ("usize", "u64", "u32"):
            if isinstance(v, float):
                if v != v:
                    return 0
                return max(0, min(int(v), M.U64 - 1))      # float -> int casts saturate
            return M.wrap_usize(int(v))
        if ty in ("i64", "i32", "isize"):
            if isinstance(v, float):
                return 0 if v != v else int(v)
            v = int(v)
            return v - M.U64 if v >= (1 << 63) else v
        raise RustRuntimeError(f"cast to {ty}")

    def ev_try(self, n, env):
        v = self.ev(n[1], env)
        if isinstance(v, Ok) or isinstance(v, Some):
            return v.v
        if isinstance(v, Err) or v is None:
            raise _Return(v)
        raise RustRuntimeError(f"line {n[2]}: `?` on {type(v).__name__}")

    def ev_tfield(self, n, env):
        return self.ev(n[1], env)[n[2]]

    def ev_field(self, n, env):
        v = self.ev(n[1], env)
        try:
            return getattr(v, n[2])
        except AttributeError:
            raise RustRuntimeError(f"line {n[3]}: no field `{n[2]}` on {type(v).__name__}")

    def ev_index(self, n, env):
        v = self.ev(n[1], env)
        i = self.ev(n[2], env)
        if isinstance(v, M.ChunkedArray):
            return v.index(i)
        if isinstance(v, list):
            if not isinstance(i, int) or i < 0 or i >= len(v):
                raise RustPanic(f"line {n[3]}: index out of bounds: the len is {len(v)} but the index is {i}")
            return v[i]
        raise RustRuntimeError(f"line {n[3]}: cannot index {type(v).__name__}")

    def ev_bin(self, n, env):
        op = n[1]
        if op == "&&":
            return bool(self.ev(n[2], env)) and bool(self.ev(n[3], env))
        if op == "||":
            return bool(self.ev(n[2], env)) or bool(self.ev(n[3], env))
        a = self.ev(n[2], env)
        b = self.ev(n[3], env)
        return self.binop(op, a, b, n[4])

    def binop(self, op, a, b, line=0):
        if isinstance(a, M.ChunkedArray) or isinstance(b, M.ChunkedArray):
            return M.chunked_binop(op, a, b)
        fa, fb = isinstance(a, float), isinstance(b, float)
        if op in ("==", "!=", "<", ">", "<=", ">="):
            if fa != fb and not (isinstance(a, bool) or isinstance(b, bool)):
                raise RustRuntimeError(f"line {line}: comparison of mixed int / float")
            return {"==": a == b, "!=": a != b, "<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b}[op]
        if fa and fb:
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if op == "/":
                return M.fdiv(a, b)
            if op == "%":
                import math
                return math.fmod(a, b) if b != 0.0 else math.nan
        if fa != fb:
            raise RustRuntimeError(f"line {line}: arithmetic on mixed int / float ({a!r} {op} {b!r})")
        # integers: usize with release-profile wrap-around
        if op == "+":
            return M.wrap_usize(a + b)
        if op == "-":
            return M.wrap_usize(a - b)
        if op == "*":
            return M.wrap_usize(a * b)
        if op == "/":
            if b == 0:
                raise RustPanic("attempt to divide by zero")
            return a // b
        if op == "%":
            if b == 0:
                raise RustPanic("attempt to calculate the remainder with a divisor of zero")
            return a % b
        raise RustRuntimeError(f"line {line}: operator {op}")

    def ev_assign(self, n, env):
        op, target, rhs = n[1], n[2], self.ev(n[3], env)
        if op != "=":
            rhs = self.binop(op[0], self.ev(target, env), rhs, n[4])
        k = target[0]
        if k == "var":
            e = env.lookup(target[1])
            if e is None:
                raise RustRuntimeError(f"line {n[4]}: assignment to unknown `{target[1]}`")
            e.vars[target[1]] = rhs
        elif k == "index":
            v = self.ev(target[1], env)
            i = self.ev(target[2], env)
            if not isinstance(v, list) or i < 0 or i >= len(v):
                raise RustPanic(f"line {n[4]}: index out of bounds: the len is {len(v)} but the index is {i}")
            v[i] = rhs
        elif k == "paren":
            return self.ev_assign(("assign", "=", target[1], ("lit", rhs), n[4]), env)
        else:
            raise RustRuntimeError(f"line {n[4]}: assignment target {k}")
        return ()

    def ev_macro(self, n, env):
        if n[1] == "izip":
            cols = [self.iterate(self.ev(a, env)) for a in n[2]]
            return M.RustIter(zip(*cols))
        raise RustRuntimeError(f"line {n[3]}: macro {n[1]}!")

    def ev_call(self, n, env):
        f = self.ev(n[1], env)
        args = [self.ev(a, env) for a in n[2]]
        if not callable(f):
            raise RustRuntimeError(f"line {n[3]}: call of non-function")
        return f(*args)

    def ev_method(self, n, env):
        recv = self.ev(n[1], env)
        name = n[2]
        args = [self.ev(a, env) for a in n[3]]
        return self.method(recv, name, args, n[4])

    # ---- methods -------------------------------------------------------------------------------
    def method(self, r, name, a, line):
        if isinstance(r, float):
            if name == "abs":
                return abs(r)
            if name == "max":
                return M.fmax(r, a[0])
            if name == "min":
                return M.fmin(r, a[0])
            if name == "sqrt":
                return M.fsqrt(r)
            if name == "mul_add":
                return M.fma(r, a[0], a[1])
            if name == "powi":
                return M.powi(r, a[0])
            if name == "is_nan":
                return r != r
        elif isinstance(r, bool):
            pass
        elif isinstance(r, int):
            if name == "saturating_sub":
                return max(r - a[0], 0)
            if name == "max":
                return max(r, a[0])
            if name == "min":
                return min(r, a[0])
            if name == "pow":
                return M.wrap_usize(r ** a[0])
        elif isinstance(r, str):
            if name in ("into", "to_string", "as_str"):
                return r
        elif isinstance(r, (Some, type(None))):
            if name == "unwrap":
                if r is None:
                    raise RustPanic(f"line {line}: called `Option::unwrap()` on a `None` value")
                return r.v
            if name == "unwrap_or":
                return a[0] if r is None else r.v
            if name == "and_then":
                return None if r is None else a[0](r.v)
            if name == "map":
                return None if r is None else Some(a[0](r.v))
            if name == "is_some":
                return r is not None
            if name == "is_none":
                return r is None
            if name == "ok_or":
                return Err(a[0]) if r is None else Ok(r.v)
            if name in ("copied", "cloned"):
                return r
        elif isinstance(r, (Ok, Err)):
            if name == "ok":
                return Some(r.v) if isinstance(r, Ok) else None
            if name == "unwrap":
                if isinstance(r, Err):
                    raise RustPanic(f"line {line}: called `Result::unwrap()` on an `Err` value: {r.v}")
                return r.v
            if name == "is_ok":
                return isinstance(r, Ok)
        elif isinstance(r, list):
            if name == "len":
                return len(r)
            if name == "get":
                i = a[0]
                return Some(r[i]) if 0 <= i < len(r) else None
            if name in ("iter", "into_iter"):
                return M.RustIter(r)
            if name == "push":
                r.append(a[0])
                return ()
            if name in ("clone", "to_vec"):
                return list(r)
            if name == "as_slice":
                return r
            if name == "is_empty":
                return not r
            if name == "first":
                return Some(r[0]) if r else None
            if name == "last":
                return Some(r[-1]) if r else None
        elif isinstance(r, (M.RustIter, range)):
            if name == "map":
                f = a[0]
                return M.RustIter(f(x) for x in r)
            if name == "zip":
                return M.RustIter(zip(r, self.iterate(a[0])))
            if name == "enumerate":
                return M.RustIter(enumerate(r))
            if name == "for_each":
                f = a[0]
                for x in r:
                    f(x)
                return ()
            if name == "collect":
                return list(r)
            if name in ("iter", "into_iter", "copied", "cloned"):
                return r if isinstance(r, M.RustIter) else M.RustIter(r)
            if name == "rev":
                return M.RustIter(reversed(list(r)))
            if name == "sum":
                items = list(r)
                acc = 0.0 if (items and isinstance(items[0], float)) else 0
                for x in items:
                    acc = self.binop("+", acc, x)
                return acc
        elif isinstance(r, tuple):
            pass
        # model objects: a real Python method of that name
        if isinstance(r, M.PrimArray) and name == "iter":
            return M.RustIter(r.opt_items())
        f = getattr(r, name, None)
        if f is not None and callable(f) and not isinstance(r, (int, float, str, list, tuple)):
            return f(*a)
        raise RustRuntimeError(f"line {line}: no method `{name}` on {type(r).__name__}")


def _series_new(name, data):
    if isinstance(data, M.ChunkedArray):          # D2: a calc_ema / calc_sma result used as Vec<Option<f64>>
        return M.Series(M.ChunkedArray(name, data.chunks, data.dtype))
    items = list(data)
    if items and all(isinstance(x, (Some, type(None))) for x in items):
        return M.Series(M.ChunkedArray.from_options(name, items))
    if any(isinstance(x, (Some, type(None))) for x in items):
        raise RustRuntimeError("Series::new: mixed Option / value items")
    return M.Series(M.ChunkedArray.from_values(name, items))


def _struct_from_series(name, n, fields):
    fields = list(fields)
    for f in fields:
        if f.len() != n:
            return Err("ShapeMismatch: struct fields of unequal length")
    return Ok(M.StructChunked(name, fields))


def _int32_from_slice(name, data):
    return M.ChunkedArray(name, [M.PrimArray(list(data), None)], "Int32")


def _vec_with_capacity(_n=0):
    return []


CTORS = {"Some": Some, "Ok": Ok, "Err": Err}

PATHS = {
    "f64::MIN": -M.F64_MAX, "f64::MAX": M.F64_MAX, "f64::NAN": float("nan"),
    "f64::INFINITY": float("inf"), "f64::NEG_INFINITY": float("-inf"), "f64::EPSILON": 2.220446049250313e-16,
    "usize::MAX": M.U64 - 1,
    "DataType::Float64": "Float64", "DataType::Int64": "Int64", "DataType::Int32": "Int32",
    "Float64Chunked::full_null": lambda name, n: M.ChunkedArray.full_null(name, n),
    "PrimitiveChunkedBuilder::new": M.Builder,
    "VecDeque::with_capacity": lambda _n=0: M.VecDeque(),
    "VecDeque::new": lambda: M.VecDeque(),
    "ArrayVec::from": lambda items: M.ArrayVec(items),
    "Vec::with_capacity": _vec_with_capacity, "Vec::new": _vec_with_capacity,
    "Series::new": _series_new,
    "StructChunked::from_series": _struct_from_series,
    "Int32Chunked::from_slice": _int32_from_slice,
    "Field::new": lambda name, dtype: ("Field", name, dtype),
    "DataType::Struct": lambda fields: ("Struct", fields),
}
