"""The CUDA path, called through the reference-facing API (polars_quant_b200.talib -> the `_polars_plugin_*` symbols
of libpqb200.so), against the vectors made by EXECUTING the reference's own text (tests/golden/talib_ref_golden.npz):
every value bit-identical, every validity bit identical, failures where the reference fails.

Inputs carry exactly the golden case's physical shape where the boundary can express it (chunks, nulls).  One
deliberate canonicalisation: a validity bitmap with every bit set is the same logical column as
no bitmap, so the product follows the reference's no-bitmap branch for it; the three reference functions whose two
branches differ on dense data (midprice overlap.rs:363 vs :384, calc_dema :560 vs :603, calc_t3 :1058 vs :1160) are
compared on the no-bitmap cases only."""
import collections

import numpy as np
import pytest

import refgolden

pa = pytest.importorskip("pyarrow")
pytestmark = pytest.mark.gpu

TAGS = ["A", "B", "Bs", "C3", "Cb", "D", "E0", "E1", "E2", "E10", "F", "G"]

# functions of the reference that are defective there and not built on the GPU (SURVEY.md 8a "oracle-only")
NOT_BUILT = {"wma", "dema", "t3", "kama"}
NOT_BUILT_MATYPES = {2, 3, 6, 8}


def _arrow(cols, name, entry):
    v, ok = cols[name]
    if ok is None or entry.get("force_bitmap"):
        arr = pa.array(v, type=pa.float64())
        if entry.get("force_bitmap") and len(v):
            bitmap = pa.py_buffer(np.packbits(np.ones(len(v), np.uint8), bitorder="little").tobytes())
            arr = pa.Array.from_buffers(pa.float64(), len(v), [bitmap, arr.buffers()[1]], null_count=0)
    else:
        arr = pa.array(v, type=pa.float64(), mask=~ok)
    chunks = entry.get("chunks")
    if chunks:
        parts, a = [], 0
        for c in chunks:
            parts.append(arr.slice(a, c))
            a += c
        return pa.chunked_array(parts)
    return arr


def _np(out):
    arr = out.combine_chunks() if isinstance(out, pa.ChunkedArray) else out
    ok = ~np.asarray(arr.is_null().to_numpy(zero_copy_only=False))
    vals = np.asarray(arr.to_numpy(zero_copy_only=False), dtype=np.float64)
    return vals, ok


def gpu_call(entry, cols):
    """-> list of pyarrow arrays, or None when the product has no such function (counted by the caller)."""
    from polars_quant_b200 import talib as T

    fn, kw, pr = entry["fn"], entry.get("kwargs", {}), entry["params"]
    a = lambda name: _arrow(cols, name, entry)
    if entry["kind"] == "py":
        if fn == "STOCH":
            return list(T.STOCH(a("high"), a("low"), a("close"), *pr))
        if fn == "MACDFIX":
            return list(T.MACDFIX(a("close"), *pr))
        f = getattr(T, fn, None)
        if f is None:
            return None
        if fn == "STOCHF":
            return list(f(a("high"), a("low"), a("close"), *pr))
        return list(f(a("close"), *pr))
    upper = fn.upper()
    if fn in ("sma", "ema", "tema", "trima", "midpoint"):
        return [getattr(T, upper)(a("close"), kw.get("timeperiod", 14 if fn == "midpoint" else 30))]
    if fn == "ma":
        return [T.MA(a("close"), kw.get("timeperiod", 30), kw.get("matype", 0))]
    if fn == "bbands":
        return list(T.BBANDS(a("close"), kw.get("timeperiod", 20), kw.get("nbdevup", 2.0), kw.get("nbdevdn", 2.0)))
    if fn == "midprice":
        return [T.MIDPRICE(a("high"), a("low"), kw.get("timeperiod", 14))]
    if fn in ("atr", "natr"):
        return [getattr(T, upper)(a("high"), a("low"), a("close"), kw.get("timeperiod", 14))]
    if fn == "trange":
        return [T.TRANGE(a("high"), a("low"), a("close"))]
    if fn == "obv":
        return [T.OBV(a("close"), a("volume"))]
    if fn == "ad":
        return [T.AD(a("high"), a("low"), a("close"), a("volume"))]
    if fn == "adosc":
        return [T.ADOSC(a("high"), a("low"), a("close"), a("volume"), kw.get("fastperiod", 3), kw.get("slowperiod", 10))]
    if fn in ("rsi", "mom", "roc", "rocp", "rocr", "rocr100", "cmo", "trix"):
        return [getattr(T, upper)(a("close"), *pr)]
    if fn == "macd":
        return list(T.MACD(a("close"), *pr))
    if fn in ("willr", "cci", "adx", "adxr", "dx", "plus_di", "minus_di", "ultosc"):
        return [getattr(T, upper)(a("high"), a("low"), a("close"), *pr)]
    if fn in ("plus_dm", "minus_dm"):
        return [getattr(T, upper)(a("high"), a("low"), *pr)]
    if fn == "aroon":
        return list(T.AROON(a("high"), a("low"), *pr))
    if fn == "mfi":
        return [T.MFI(a("high"), a("low"), a("close"), a("volume"), *pr)]
    return None


def _windows_without_nan(cols, entry):
    """bool per bar: none of the function's inputs holds a NaN value in [t - p + 1, t]."""
    p = int(entry["kwargs"].get("timeperiod", 14))
    n = len(cols[entry["cols"][0]][0])
    bad = np.zeros(n, bool)
    for c in entry["cols"]:
        bad |= np.isnan(cols[c][0])
    hit = np.zeros(n, bool)
    for t in np.flatnonzero(bad):
        hit[t:t + max(p, 1)] = True
    return ~hit


def _branch_dependent(entry):
    return entry.get("force_bitmap") and (entry["fn"] in ("midprice", "dema", "t3") or
                                          (entry["fn"] == "ma" and entry["kwargs"].get("matype") in (3, 8)))


@pytest.mark.parametrize("tag", TAGS)
def test_gpu_reproduces_the_executed_reference(tag):
    from polars_quant_b200.plugin import PluginError

    g, index = refgolden.load()
    done = collections.Counter()
    missing, failures = [], []
    for e in index:
        if e["tag"] != tag or _branch_dependent(e):
            continue
        if e["fn"] in ("bop", "avgprice", "medprice", "typprice", "wclprice"):
            continue                                    # the candle engine's outputs: tests/test_gpu_candles.py
        name = f"{tag}/{e['key']}"
        if e["fn"] in NOT_BUILT:
            done["defective_in_the_reference_not_built"] += 1      # no product function of that name (oracle-only)
            continue
        not_built = e["fn"] == "ma" and e["kwargs"].get("matype") in NOT_BUILT_MATYPES
        cols = refgolden.inputs(g, e)
        want = refgolden.expected(g, e)
        try:
            got = gpu_call(e, cols)
        except PluginError as err:
            if not_built:
                assert "UNSUPPORTED" in str(err) or "not built" in str(err), f"{name}: {err}"
                done["refused_not_built"] += 1
                continue
            if e["fn"] == "midprice" and want is None:
                done["fails_alike"] += 1
                continue
            if want is not None and e["fn"] in ("midpoint", "adosc") and ("UNSUPPORTED" in str(err) or "leading nulls only" in str(err)):
                # null-skipping in the reference; the optional groups are built for leading nulls only
                done["interior_nulls_not_built"] += 1
                continue
            if want is not None:
                failures.append(f"{name}: the GPU path fails ({err}) where the reference succeeds")
                continue
            done["fails_alike"] += 1
            continue
        if got is None:
            done["no_entry_point"] += 1
            missing.append(e["key"])
            continue
        assert not not_built or len(cols["close"][0]) == 0, f"{name}: expected a refusal"
        if want is None:
            # The reference aborts or errors here.  The product may answer instead of aborting only for midprice
            # with nulls in `low` (the reference dies in polars' arithmetic; the product returns all-null).
            if e["fn"] != "midprice":
                failures.append(f"{name}: the reference fails (err {e['err']}) where the GPU path answers")
            done["answered_where_reference_aborts"] += 1
            continue
        assert len(got) == len(want), name
        bad = False
        for j, (arr, (gv, gok)) in enumerate(zip(got, want)):
            vals, ok = _np(arr)
            if e["fn"] in ("midprice", "midpoint"):
                # monotonic deques with NaN VALUES: container-order artifacts of the reference; bars
                # whose window holds a NaN are left out, every other bar must be exact
                keep = _windows_without_nan(cols, e)
                vals, ok, gv, gok = vals[keep], ok[keep], gv[keep], gok[keep]
            msg = refgolden.same(vals, ok, gv, gok)
            if msg:
                failures.append(f"{name}/{j}: {msg}")
                bad = True
        done["differs" if bad else "bit_exact"] += 1
    print(tag, dict(done))
    assert not failures, "\n".join(failures[:40])
    assert done["bit_exact"] + done["fails_alike"] >= 60, done
    assert not missing, missing                        # every golden call has a product entry point
