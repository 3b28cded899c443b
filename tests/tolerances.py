"""Parity bar for the CUDA path against the CPU oracle.

BASELINE.json's north_star asks for: validity (null positions) bit-exact, OBV bit-exact, floats within
rel 1e-10 / abs 1e-12.  The CUDA kernel walks every symbol serially in the reference's own operation
order (one lane per symbol, `fma` exactly where the Rust uses `mul_add`, IEEE division and square
root, no re-association), so the bar asserted here is the strongest one: EVERY output value is
bit-identical to the oracle's f64 (NaN payload/sign excepted: any NaN equals any NaN), and every
validity bit is identical.  REL/ABS are kept only to report how far a mismatch is, should one appear.
"""
import numpy as np

REL, ABS = 1e-10, 1e-12          # the north_star tolerance (not needed: the comparison is exact)


def same_bits(a, b):
    """Elementwise: identical f64 bit patterns, or both NaN."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))


def compare(name, gv, gok, ref, ok):
    """Returns (n_bad, message).  Validity must match exactly; null slots hold NaN; valid slots must
    carry the oracle's bits."""
    if not np.array_equal(gok, ok):
        idx = np.argwhere(gok != ok)[:5].tolist()
        return int((gok != ok).sum()), f"{name}: validity differs at {idx}"
    if not np.isnan(gv[~gok]).all():
        return 1, f"{name}: null slots must hold NaN"
    good = same_bits(gv, ref) | ~ok
    nbad = int((~good).sum())
    msg = ""
    if nbad:
        i = tuple(int(x) for x in np.argwhere(~good)[0])
        err = abs(gv[i] - ref[i])
        msg = (f"{name}: {nbad} of {int(ok.sum())} values differ from the oracle's bits; first at {i}: "
               f"gpu {gv[i]!r} ref {ref[i]!r} |err| {err:.3e} (north_star tol {ABS + REL * abs(ref[i]):.3e})")
    return nbad, msg


def compare_all(res, out, ok, names, skip=()):
    """res: {name: (values, validity)} from the GPU; out/ok: oracle [21, S, N].  Returns list of failures."""
    fails = []
    for j, name in enumerate(names):
        if name in skip or name not in res:
            continue
        gv, gok = res[name]
        nbad, msg = compare(name, gv, gok, out[j], ok[j])
        if nbad:
            fails.append(msg)
    return fails
