"""The batched column boundary (pqb_panel_set_columns / pqb_suite_run_columns / pqb_panel_export_arrow), the per-block
null dispatch behind it, and the headline launch shape (50,000 x 5,040, device-resident) against the oracle."""
import ctypes as C

import numpy as np
import pytest

import devread
import synth
import tolerances as T
from oracle import pqo

pytestmark = pytest.mark.gpu
F = ("close", "high", "low", "volume")


@pytest.fixture(scope="module")
def pq():
    import polars_quant_b200 as pq
    return pq


def test_run_columns_from_caller_buffers_equals_the_staged_path(pq):
    S, N = 300, 500
    d = synth.ohlcv(S, N, seed=77)
    a = pq.Panel(S, N)
    a.set_fields(d["close"], d["high"], d["low"], d["volume"])
    ra = {k: (v.copy(), ok.copy()) for k, (v, ok) in a.compute().items()}
    b = pq.Panel(S, N)
    refs, keep = pq.Panel.field_refs(d["close"], d["high"], d["low"], d["volume"])
    for threads in (1, 4):
        b.run_columns(refs, threads=threads)
        rb = b.outputs()
        for name in ra:
            assert np.array_equal(ra[name][1], rb[name][1]), name
            assert T.same_bits(ra[name][0], rb[name][0]).all(), name
    out, ok, _ = pqo.suite_panel(d["close"], d["high"], d["low"], d["volume"])
    assert not T.compare_all(b.outputs(), out, ok, pqo.OUTPUT_NAMES)
    a.close(); b.close()


def test_batched_intake_with_nulls_and_per_block_dispatch(pq):
    """Six symbol blocks, interior nulls only in blocks 1 and 4: those run the null-aware kernel, the rest the plain one
    (two launches over block lists).  Every symbol must equal what it gives alone."""
    S, N = 190, 420
    d = synth.ohlcv(S, N, seed=5)
    ok = {f: np.ones((S, N), bool) for f in F}
    for f in F:
        ok[f][3, :50] = False                    # block 0: a later listing (leading nulls, all fields): plain kernel
        ok[f][100, :7] = False                   # block 3: same
    ok["close"][40, 200:203] = False             # block 1: a halt
    ok["volume"][45, :30] = False                # block 1: a field that starts later
    for f in F:
        ok[f][150, N - 25:] = False              # block 4: delisted
    cols = []
    for s in range(S):
        for f in F:
            bits = None if ok[f][s].all() else np.packbits(ok[f][s].astype(np.uint8), bitorder="little")
            cols.append((s, f, d[f][s], bits))
    refs, keep = pq.Panel.col_refs(cols)
    p = pq.Panel(S, N)
    p.run_columns(refs, threads=3)
    got = {k: (v.copy(), o.copy()) for k, (v, o) in p.outputs().items()}
    # the same columns one by one through pqb_panel_set_column + compute()
    q = pq.Panel(S, N)
    for s, f, v, bits in cols:
        q.set_column(s, f, v, validity=bits)
    want = q.compute()
    for name in want:
        assert np.array_equal(want[name][1], got[name][1]), name
        assert T.same_bits(want[name][0], got[name][0]).all(), name
    # plain blocks are the oracle's dense computation from each symbol's first valid bar
    for s in (0, 3, 70, 100, 189):
        a = int(np.argmax(ok["close"][s]))
        out, okk, _ = pqo.suite_panel(*(d[f][s:s + 1, a:] for f in F))
        for j, name in enumerate(pqo.OUTPUT_NAMES):
            fv = np.full(N, np.nan); fk = np.zeros(N, bool)
            fv[a:], fk[a:] = out[j, 0], okk[j, 0]
            nbad, msg = T.compare(name, got[name][0][s], got[name][1][s], fv, fk)
            assert nbad == 0, f"symbol {s}: {msg}"
    # a halted symbol: null-skipping functions against the oracle's null semantics
    for name, ref in (("sma", pqo.sma(d["close"][40], 30, ok["close"][40])), ("ema", pqo.ema(d["close"][40], 30, ok["close"][40])),
                      ("obv", pqo.obv(d["close"][45], d["volume"][45], ok["close"][45], ok["volume"][45]))):
        s = 40 if name != "obv" else 45
        nbad, msg = T.compare(name, got[name][0][s], got[name][1][s], ref[0], ref[1])
        assert nbad == 0, msg
    assert not got["macd"][1][40].any() and not got["rsi"][1][40].any()        # momentum.rs refuses nulls
    # overwriting the halted columns with null-free ones returns the blocks to the plain kernel (nothing sticky)
    fix = [(s, f, d[f][s], None) for s in (40, 45, 150) for f in F]
    refs2, keep2 = pq.Panel.col_refs(fix)
    p.set_columns(refs2)
    res = p.compute()
    out, okk, _ = pqo.suite_panel(*(d[f][40:41] for f in F))
    for j, name in enumerate(pqo.OUTPUT_NAMES):
        nbad, msg = T.compare(name, res[name][0][40], res[name][1][40], out[j, 0], okk[j, 0])
        assert nbad == 0, msg
    p.close(); q.close()


def test_export_arrow_is_one_zero_copy_record_batch(pq):
    pa = pytest.importorskip("pyarrow")
    S, N = 70, 333
    d = synth.ohlcv(S, N, seed=9)
    p = pq.Panel(S, N)
    refs, keep = pq.Panel.field_refs(d["close"], d["high"], d["low"], d["volume"])
    p.run_columns(refs)
    names = ["SYM%03d" % s for s in range(S)]
    rb = p.export_arrow(symbol_names=names)
    assert rb.num_columns == S * 21 and rb.num_rows == N
    assert rb.schema.names[:3] == ["SYM000_sma", "SYM000_ema", "SYM000_tema"] and rb.schema.names[-1] == "SYM069_midprice"
    res = p.outputs()
    for s in (0, 31, 32, 69):
        for k, name in enumerate(pqo.OUTPUT_NAMES):
            col = rb.column(s * 21 + k)
            okk = ~np.asarray(col.is_null())
            assert np.array_equal(okk, res[name][1][s]), (s, name)
            v = col.to_numpy(zero_copy_only=False)
            assert T.same_bits(v[okk], res[name][0][s][okk]).all(), (s, name)
    # the batch keeps the native panel (its pinned planes) alive after the Python object is gone
    host_ptr = pq._native.lib().pqb_panel_host_output(p._h, 0)
    assert rb.column(0).buffers()[1].address == host_ptr
    want = rb.column(5).to_numpy(zero_copy_only=False).copy()
    p.close()
    del p
    q = pq.Panel(S, N)                                       # would take the pinned planes from the pool if they had been freed
    q.set_fields(d["close"][::-1].copy(), d["high"], d["low"], d["volume"])
    q.compute()
    assert T.same_bits(rb.column(5).to_numpy(zero_copy_only=False), want).all()
    q.close()


def test_headline_launch_shape_against_the_oracle(pq):
    """BASELINE config 4 as bench.py times it: 50,000 x 5,040 device-resident, one launch of 1,563 CTAs (3.5 waves of three
    CTAs per SM).  Symbol blocks of every wave, the wave boundaries and the ragged last block are read back from HBM and
    compared with the oracle bit for bit (values and validity)."""
    S, N = 50_000, 5_040
    lib = pq._native.lib()
    p = pq.Panel(S, N, host_staging=False)
    p.fill_synthetic(seed=0xC0FFEE, sigma=0.02)
    p.run()
    p.sync()
    nb, bp = p.tiled_shape()
    assert nb == 1563
    blocks = [0, 1, 147, 148, 295, 443, 444, 445, 887, 888, 1000, 1331, 1332, 1479, 1500, 1561, 1562]
    for b in blocks:
        s0, ns = b * 32, min(32, S - b * 32)
        fin = [devread.read_block(lib.pqb_panel_device_field(p._h, f), b, bp, N)[:ns] for f in range(4)]
        out, ok, _ = pqo.suite_panel(fin[0], fin[1], fin[2], fin[3])
        for k, name in enumerate(pqo.OUTPUT_NAMES):
            gv = devread.read_block(lib.pqb_panel_device_output(p._h, k), b, bp, N)[:ns]
            gok = devread.read_validity_rows(lib.pqb_panel_device_validity(p._h, k), s0, ns, p.validity_pitch, N)
            nbad, msg = T.compare(name, gv, gok, out[k], ok[k])
            assert nbad == 0, f"block {b}: {msg}"
    p.close()
