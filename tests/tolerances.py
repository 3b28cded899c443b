"""Parity tolerances for the CUDA path against the CPU oracle (see DESIGN.md "numerics").

Bar (BASELINE.json north_star): validity bit-exact; OBV bit-exact; floats |gpu - ref| <=
ABS + REL*|ref| with REL = 1e-10, ABS = 1e-12.

Five outputs are *differences of much larger quantities*, and the reference's own value there
carries rounding noise well above ABS, because its sliding sums are running sums
(sum += new; sum -= old: a random walk of roundings of size ~eps*|sum|*sqrt(2t)) and its EMAs are
only accurate to a few eps*price.  For those outputs -- and only those -- an extra absolute
allowance `cond` proportional to eps times the magnitude of the operands is added; it is the level
at which two correct evaluation orders of the reference's own formula already disagree:

  macd, macd_signal, macd_hist : fast EMA - slow EMA (and its EMA)     cond = 64*eps*|close|
  kdj_j = 3K - 2D (K, D are running-sum SMAs of fastk)                cond = 4*sqrt(n)*eps*(3|K|+2|D|)
  bb_upper / bb_lower : mean +- nbdev*sqrt(sum_sq/p - mean^2)          cond = nbdev*4*sqrt(n)*eps*mean^2/(2*sd)
  ad : sign-indefinite running sum                                     cond = 4*sqrt(n)*eps*max_{s<=t}|ad_s|
"""
import numpy as np

REL, ABS = 1e-10, 1e-12
EPS = np.finfo(np.float64).eps


def tolerance(name, ref, ok, ctx):
    """ref: oracle values [.., n_bars]; ok: its validity; ctx: dict(close=..., out={name: values}, nbdevup, nbdevdn)."""
    n = ref.shape[-1]
    tol = ABS + REL * np.abs(np.nan_to_num(ref))
    drift = 4.0 * np.sqrt(max(n, 1)) * EPS
    with np.errstate(divide="ignore", invalid="ignore"):
        if name in ("macd", "macd_signal", "macd_hist"):
            tol = tol + 64.0 * EPS * np.abs(ctx["close"])
        elif name == "kdj_j":
            k, d = np.nan_to_num(ctx["out"]["kdj_k"]), np.nan_to_num(ctx["out"]["kdj_d"])
            tol = tol + drift * (3.0 * np.abs(k) + 2.0 * np.abs(d))
        elif name in ("bb_upper", "bb_lower"):
            mid = np.nan_to_num(ctx["out"]["bb_middle"])
            nb = ctx.get("nbdevup", 2.0) if name == "bb_upper" else ctx.get("nbdevdn", 2.0)
            sd = np.abs(np.nan_to_num(ref) - mid) / max(abs(nb), 1e-300)
            cond = abs(nb) * drift * mid * mid / (2.0 * np.maximum(sd, 1e-300))
            # sd == 0 in the reference means the (noisy) variance clipped at 0: anything up to the
            # square root of the variance noise is equally valid
            cond = np.where(sd > 0, cond, abs(nb) * np.sqrt(drift) * np.abs(mid))
            tol = tol + np.minimum(cond, abs(nb) * np.sqrt(drift) * np.abs(mid) + 1e-300)
        elif name == "ad":
            tol = ABS + drift * np.maximum.accumulate(np.abs(np.nan_to_num(ref)), axis=-1)
    return tol


def compare(name, gv, gok, ref, ok, ctx):
    """Returns (n_bad, worst err/tol ratio, message).  Validity must match exactly; null slots hold NaN."""
    if not np.array_equal(gok, ok):
        idx = np.argwhere(gok != ok)[:5].tolist()
        return int((gok != ok).sum()), np.inf, f"{name}: validity differs at {idx}"
    if not np.isnan(gv[~gok]).all():
        return 1, np.inf, f"{name}: null slots must hold NaN"
    if name == "obv":
        bad = gv[ok] != ref[ok]
        return int(bad.sum()), (np.inf if bad.any() else 0.0), f"obv: {int(bad.sum())} values not bit-exact"
    tol = tolerance(name, ref, ok, ctx)
    err = np.abs(gv - ref)
    m = ok & ~(np.isnan(ref) & np.isnan(gv))
    ratio = np.zeros_like(err)
    ratio[m] = err[m] / tol[m]
    ratio[m & np.isnan(ratio)] = np.inf
    nbad = int((ratio > 1.0).sum())
    worst = float(ratio.max()) if ratio.size else 0.0
    msg = ""
    if nbad:
        i = np.unravel_index(np.argmax(ratio), ratio.shape)
        msg = f"{name}: {nbad} of {int(m.sum())} outside tol; worst at {i}: gpu {gv[i]!r} ref {ref[i]!r} err {err[i]:.3e} tol {tol[i]:.3e}"
    return nbad, worst, msg


def compare_all(res, out, ok, close, names, nbdevup=2.0, nbdevdn=2.0, skip=()):
    """res: {name: (values, validity)} from the GPU; out/ok: oracle [21, S, N].  Returns (failures, worst-by-name)."""
    ctx = {"close": close, "out": {n: out[j] for j, n in enumerate(names)}, "nbdevup": nbdevup, "nbdevdn": nbdevdn}
    fails, worst = [], {}
    for j, name in enumerate(names):
        if name in skip:
            continue
        gv, gok = res[name]
        nbad, w, msg = compare(name, gv, gok, out[j], ok[j], ctx)
        worst[name] = w
        if nbad:
            fails.append(msg)
    return fails, worst
