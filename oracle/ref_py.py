"""Independent pure-Python restatement of the reference's src/talib loops.

TEST INFRASTRUCTURE ONLY (see oracle/pq_oracle.c header).  Written separately from the C
oracle, straight from the Rust text, in the Rust's own vocabulary (`Option` -> None,
`VecDeque` -> collections.deque, `mul_add` -> an exactly rounded fma built on Fraction) so
that two restatements of the same source can be compared bit-for-bit on small inputs, and
so that golden vectors (tests/golden/) exist that were not produced by the C oracle.

PARITY UNPINNED: the reference has no golden vectors for this path and cannot run here.

A column is a Python list of `float | None`.  Errors the reference would raise are
`RefError` (PolarsResult::Err) -- e.g. momentum.rs functions call cont_slice()? which
fails on nulls.
"""
from __future__ import annotations

import math
from collections import deque
from fractions import Fraction

F64_MIN = -1.7976931348623157e308
F64_MAX = 1.7976931348623157e308
U64 = 1 << 64


class RefError(Exception):
    pass


def fma(a: float, b: float, c: float) -> float:
    """f64::mul_add: a*b+c with a single rounding."""
    if not (math.isfinite(a) and math.isfinite(b) and math.isfinite(c)):
        return a * b + c
    r = Fraction(a) * Fraction(b) + Fraction(c)
    if r == 0:
        # sign of an exact zero sum follows IEEE: +0 unless both addends are -0
        return a * b + c
    try:
        return float(r)
    except OverflowError:
        return math.copysign(math.inf, r)


def rs_max(a: float, b: float) -> float:  # f64::max ignores a NaN operand
    if math.isnan(a):
        return b
    if math.isnan(b):
        return a
    return a if a > b else b


def rs_min(a: float, b: float) -> float:
    if math.isnan(a):
        return b
    if math.isnan(b):
        return a
    return a if a < b else b


def _cont_slice(*cols):
    for col in cols:
        if any(v is None for v in col):
            raise RefError("chunked array is not contiguous")


# ---------------------------------------------------------------- overlap.rs
def calc_sma(values, timeperiod):  # overlap.rs:871-937
    n = len(values)
    if timeperiod == 0 or n < timeperiod:
        return [None] * n
    out = []
    denominator = 1.0 / float(timeperiod)
    count, s, window = 0, 0.0, deque()
    for value in values:
        if value is None:
            out.append(None)
            continue
        count += 1
        s += value
        window.append(value)
        if count < timeperiod:
            out.append(None)
        else:
            if count > timeperiod:
                old = window.popleft()
                s -= old
                count -= 1
            out.append(s * denominator)
    return out


def calc_ema(values, timeperiod):  # overlap.rs:660-730
    n = len(values)
    if timeperiod == 0 or n < timeperiod:
        return [None] * n
    alpha = 2.0 / (float(timeperiod) + 1.0)
    out, count, ema_value, s = [], 0, 0.0, 0.0
    for value in values:
        if value is None:
            out.append(None)
            continue
        count += 1
        if count < timeperiod:
            s += value
            out.append(None)
        elif count == timeperiod:
            s += value
            ema_value = s / float(timeperiod)
            out.append(ema_value)
        else:
            ema_value = fma(alpha, value - ema_value, ema_value)
            out.append(ema_value)
    return out


def calc_rma(x, timeperiod):  # D1 (frozen; see pq_oracle.c header)
    n = len(x)
    if timeperiod == 0 or n < timeperiod:
        return [None] * n
    a = 1.0 / float(timeperiod)
    out, s, y = [], 0.0, 0.0
    for i, v in enumerate(x):
        if i < timeperiod - 1:
            s += v
            out.append(None)
        elif i == timeperiod - 1:
            s += v
            y = s / float(timeperiod)
            out.append(y)
        else:
            y = fma(a, v - y, y)
            out.append(y)
    return out


def calc_tema(values, timeperiod):  # overlap.rs:1177-1311
    n = len(values)
    p = timeperiod
    if p == 0 or n < 3 * p - 2:
        return [None] * n
    alpha = 2.0 / (float(p) + 1.0)
    e, s, count, out = [0.0] * 3, [0.0] * 3, 0, []
    for value in values:
        if value is None:
            out.append(None)
            continue
        count += 1
        if count < p:
            s[0] += value
            out.append(None)
        elif count == p:
            s[0] += value
            e[0] = s[0] / float(p)
            s[1] = e[0]
            out.append(None)
        elif count < 2 * p - 1:
            e[0] = fma(alpha, value - e[0], e[0])
            s[1] += e[0]
            out.append(None)
        elif count == 2 * p - 1:
            e[0] = fma(alpha, value - e[0], e[0])
            s[1] += e[0]
            e[1] = s[1] / float(p)
            s[2] = e[1]
            out.append(None)
        elif count < 3 * p - 2:
            e[0] = fma(alpha, value - e[0], e[0])
            e[1] = fma(alpha, e[0] - e[1], e[1])
            s[2] += e[1]
            out.append(None)
        elif count == 3 * p - 2:
            e[0] = fma(alpha, value - e[0], e[0])
            e[1] = fma(alpha, e[0] - e[1], e[1])
            s[2] += e[1]
            e[2] = s[2] / float(p)
            out.append(3.0 * e[0] - 3.0 * e[1] + e[2])
        else:
            e[0] = fma(alpha, value - e[0], e[0])
            e[1] = fma(alpha, e[0] - e[1], e[1])
            e[2] = fma(alpha, e[1] - e[2], e[2])
            out.append(3.0 * e[0] - 3.0 * e[1] + e[2])
    return out


def calc_trima(values, timeperiod):  # overlap.rs:1313-1326
    if timeperiod % 2 == 1:
        n = timeperiod // 2 + 1
        return calc_sma(calc_sma(values, n), n)
    n = timeperiod // 2
    return calc_sma(calc_sma(values, n), n + 1)


def calc_wma(values, timeperiod):  # overlap.rs:1328-1399 (literal)
    n = len(values)
    if timeperiod == 0 or n < timeperiod:
        return [None] * n
    denominator = float(timeperiod * (timeperiod + 1) // 2)
    out, count, numerator, window = [], 0, 0.0, deque()
    for value in values:
        if value is None:
            out.append(None)
            continue
        count += 1
        numerator += float(count) * value
        window.append(value)
        if count < timeperiod:
            out.append(None)
        else:
            if count > timeperiod:
                old = window.popleft()
                numerator -= float(timeperiod) * old
                count -= 1
            out.append(numerator / denominator)
    return out


def calc_ma(values, timeperiod, matype):  # overlap.rs:857-869
    if matype == 1:
        return calc_ema(values, timeperiod)
    if matype == 2:
        return calc_wma(values, timeperiod)
    if matype in (3, 6, 8):
        raise RefError("matype %d not restated (reference defect, SURVEY 8a)" % matype)
    if matype == 4:
        return calc_tema(values, timeperiod)
    if matype == 5:
        return calc_trima(values, timeperiod)
    return calc_sma(values, timeperiod)


def bbands(real, timeperiod=20, nbdevup=2.0, nbdevdn=2.0):  # overlap.rs:47-116
    n = len(real)
    if timeperiod == 0 or n < timeperiod:
        return [None] * n, [None] * n, [None] * n
    up, mid, lo = [], [], []
    count, s, ss, window = 0, 0.0, 0.0, deque()
    for value in real:
        if value is None:
            up.append(None), mid.append(None), lo.append(None)
            continue
        count += 1
        s += value
        ss += value * value
        window.append(value)
        if count < timeperiod:
            up.append(None), mid.append(None), lo.append(None)
        else:
            if count > timeperiod:
                old = window.popleft()
                s -= old
                ss -= old * old
                count -= 1
            mean = s / float(timeperiod)
            variance = (ss / float(timeperiod)) - mean * mean
            std = math.sqrt(rs_max(variance, 0.0))
            up.append(mean + nbdevup * std)
            mid.append(mean)
            lo.append(mean - nbdevdn * std)
    return up, mid, lo


def midpoint(real, timeperiod=14):  # overlap.rs:180-278 (literal, incl. the :227/:264 defect)
    out, count = [], 0
    wmax, wmin = deque(), deque()
    for value in real:
        if value is None:
            out.append(None)
            continue
        count += 1
        while wmax and wmax[-1][1] <= value:
            wmax.pop()
        if wmax and wmax[0][0] == (count - timeperiod) % U64:
            wmax.popleft()
        wmax.append((count, value))
        mx = wmax[0][1]
        while wmin and wmin[-1][1] >= value:
            wmin.pop()
        if wmax and wmax[0][0] == (count - timeperiod) % U64:  # sic: tests window_max
            if wmin:
                wmin.popleft()
        wmin.append((count, value))
        mn = wmin[0][1]
        out.append((mx + mn) / 2.0)
    return out


def midprice(high, low, timeperiod=14):  # overlap.rs:281-404
    if any(v is None for v in low):
        raise RefError("midprice: low null branch builds unequal columns (overlap.rs:352-376)")
    hmax, count, w = [], 0, deque()
    for value in high:
        if value is None:
            hmax.append(None)
            continue
        count += 1
        while w and w[-1][1] <= value:
            w.pop()
        if w and w[0][0] == (count - timeperiod) % U64:
            w.popleft()
        w.append((count, value))
        hmax.append(w[0][1])
    lmin, count, w = [], 0, deque()
    for value in low:
        count += 1
        while w and w[-1][1] >= value:
            w.pop()
        if w and w[0][0] == (count - timeperiod) % U64:
            w.popleft()
        w.append((count, value))
        lmin.append(w[0][1])
    return [None if a is None else (a + b) / 2.0 for a, b in zip(hmax, lmin)]


# ---------------------------------------------------------------- momentum.rs
def rsi(real, timeperiod=14):  # momentum.rs:507-541
    _cont_slice(real)
    n = len(real)
    ups, downs = [0.0] * n, [0.0] * n
    for i in range(1, n):
        diff = real[i] - real[i - 1]
        if diff > 0.0:
            ups[i] = diff
        else:
            downs[i] = -diff
    au, ad = calc_rma(ups, timeperiod), calc_rma(downs, timeperiod)
    res = [None] * n
    for i in range(n):
        if au[i] is not None and ad[i] is not None:
            if ad[i] == 0.0:
                res[i] = 100.0
            else:
                rs = au[i] / ad[i]
                res[i] = 100.0 - (100.0 / (1.0 + rs))
    return res


def macd(real, fastperiod=12, slowperiod=26, signalperiod=9):  # momentum.rs:250-283
    _cont_slice(real)
    n = len(real)
    fast, slow = calc_ema(real, fastperiod), calc_ema(real, slowperiod)
    dif = [None] * n
    for i in range(n):
        if fast[i] is not None and slow[i] is not None:
            dif[i] = fast[i] - slow[i]
    dea = calc_ema([0.0 if v is None else v for v in dif], signalperiod)
    hist = [None] * n
    for i in range(n):
        if dif[i] is not None and dea[i] is not None:
            hist[i] = dif[i] - dea[i]
    return dif, dea, hist


def willr(high, low, close, timeperiod=14):  # momentum.rs:630-662
    _cont_slice(high, low, close)
    n = len(high)
    res = [None] * n
    if timeperiod == 0:
        return res
    for i in range(timeperiod - 1, n):
        mx, mn = F64_MIN, F64_MAX
        for j in range(i + 1 - timeperiod, i + 1):
            mx = rs_max(mx, high[j])
            mn = rs_min(mn, low[j])
        diff = mx - mn
        res[i] = 0.0 if diff == 0.0 else -100.0 * (mx - close[i]) / diff
    return res


def calc_dm(high, low, close, timeperiod):  # momentum.rs:668-727 -> (dx, minus_di)
    n = len(high)
    p_dm, m_dm, tr = [0.0] * n, [0.0] * n, [0.0] * n
    for i in range(1, n):
        up_move = high[i] - high[i - 1]
        down_move = low[i - 1] - low[i]
        if up_move > down_move and up_move > 0.0:
            p_dm[i] = up_move
        if down_move > up_move and down_move > 0.0:
            m_dm[i] = down_move
        tr[i] = rs_max(rs_max(high[i] - low[i], abs(high[i] - close[i - 1])), abs(low[i] - close[i - 1]))
    sp, sm, st = calc_rma(p_dm, timeperiod), calc_rma(m_dm, timeperiod), calc_rma(tr, timeperiod)
    plus_di, minus_di = [None] * n, [None] * n
    for i in range(n):
        if sp[i] is not None and sm[i] is not None and st[i] is not None and st[i] != 0.0:
            plus_di[i] = 100.0 * sp[i] / st[i]
            minus_di[i] = 100.0 * sm[i] / st[i]
    dx = [None] * n
    for i in range(n):
        if plus_di[i] is not None and minus_di[i] is not None:
            diff, sm_ = abs(plus_di[i] - minus_di[i]), plus_di[i] + minus_di[i]
            dx[i] = 0.0 if sm_ == 0.0 else 100.0 * diff / sm_
    return dx, minus_di


def dm_family(high, low, close, timeperiod=14):
    """plus_dm :418, minus_dm :362, plus_di :401 (returns calc_dm().0 == DX), minus_di :346, dx :226, adx :11,
    adxr :29 -- as a dict of columns."""
    _cont_slice(high, low, close)
    n = len(high)
    p_dm, m_dm = [0.0] * n, [0.0] * n
    for i in range(1, n):
        up_move = high[i] - high[i - 1]
        down_move = low[i - 1] - low[i]
        if up_move > down_move and up_move > 0.0:
            p_dm[i] = up_move
        if down_move > up_move and down_move > 0.0:
            m_dm[i] = down_move
    dx, minus_di = calc_dm(high, low, close, timeperiod)
    adx = calc_rma([0.0 if v is None else v for v in dx], timeperiod)
    adxr = [None] * n
    if timeperiod >= 1:
        for i in range(timeperiod - 1, n):
            prev = adx[max(0, i - (timeperiod - 1))]
            if adx[i] is not None and prev is not None:
                adxr[i] = (adx[i] + prev) * 0.5
    return {"plus_dm": calc_rma(p_dm, timeperiod), "minus_dm": calc_rma(m_dm, timeperiod), "plus_di": dx, "dx": dx,
            "minus_di": minus_di, "adx": adx, "adxr": adxr}


def _slice_ema(x, timeperiod):  # D2: calc_ema over a plain slice (the no-validity branch overlap.rs:705-724)
    return calc_ema(list(x), timeperiod)


def trix(real, timeperiod=30):  # momentum.rs:544-571
    _cont_slice(real)
    e1 = _slice_ema(real, timeperiod)
    e2 = _slice_ema([0.0 if v is None else v for v in e1], timeperiod)
    e3 = _slice_ema([0.0 if v is None else v for v in e2], timeperiod)
    n = len(e3)
    res = [None] * n
    for i in range(1, n):
        curr, prev = e3[i], e3[i - 1]
        if curr is not None and prev is not None and prev != 0.0:
            res[i] = (curr - prev) / prev * 100.0
    return res


def ultosc(high, low, close, p1=7, p2=14, p3=28):  # momentum.rs:573-627
    _cont_slice(high, low, close)
    n = len(high)
    bp, tr = [0.0] * n, [0.0] * n
    for i in range(1, n):
        min_l_pc = rs_min(low[i], close[i - 1])
        max_h_pc = rs_max(high[i], close[i - 1])
        bp[i] = close[i] - min_l_pc
        tr[i] = max_h_pc - min_l_pc

    def avg(p):
        res, s_bp, s_tr = [None] * n, 0.0, 0.0
        for i in range(n):
            s_bp += bp[i]
            s_tr += tr[i]
            if i >= p:
                s_bp -= bp[i - p]
                s_tr -= tr[i - p]
            if i >= p - 1 and s_tr != 0.0:
                res[i] = s_bp / s_tr
        return res

    a1, a2, a3 = avg(p1), avg(p2), avg(p3)
    return [None if (a1[i] is None or a2[i] is None or a3[i] is None) else 100.0 * (4.0 * a1[i] + 2.0 * a2[i] + a3[i]) / 7.0
            for i in range(n)]


def aroon(high, low, timeperiod=14):  # momentum.rs:63-110 -> (aroon_up, aroon_down)
    _cont_slice(high, low)
    n = len(high)
    up, down = [None] * n, [None] * n
    for i in range(timeperiod, n):
        start = i - timeperiod
        max_idx, max_val, min_idx, min_val = 0, F64_MIN, 0, F64_MAX
        for j in range(start, i + 1):
            if high[j] >= max_val:
                max_val, max_idx = high[j], j - start
            if low[j] <= min_val:
                min_val, min_idx = low[j], j - start
        up[i] = (float(max_idx) / float(timeperiod)) * 100.0
        down[i] = (float(min_idx) / float(timeperiod)) * 100.0
    return up, down


def mom(real, timeperiod=10):  # momentum.rs:384-397
    _cont_slice(real)
    n = len(real)
    res = [None] * n
    for i in range(timeperiod, n):
        res[i] = real[i] - real[i - timeperiod]
    return res


def roc(real, timeperiod=10, kind=0):  # momentum.rs:439-504; kind 0 roc 1 rocp 2 rocr 3 rocr100
    _cont_slice(real)
    n = len(real)
    res = [None] * n
    for i in range(timeperiod, n):
        curr, prev = real[i], real[i - timeperiod]
        if prev != 0.0:
            res[i] = [(curr - prev) / prev * 100.0, (curr - prev) / prev, curr / prev,
                      (curr / prev) * 100.0][kind]
    return res


def cmo(real, timeperiod=14):  # momentum.rs:181-223
    _cont_slice(real)
    n = len(real)
    ups, downs = [0.0] * n, [0.0] * n
    for i in range(1, n):
        diff = real[i] - real[i - 1]
        if diff > 0.0:
            ups[i] = diff
        else:
            downs[i] = -diff
    res, su, sd = [None] * n, 0.0, 0.0
    for i in range(n):
        su += ups[i]
        sd += downs[i]
        if i >= timeperiod:
            su -= ups[i - timeperiod]
            sd -= downs[i - timeperiod]
        if timeperiod >= 1 and i >= timeperiod - 1:
            total = su + sd
            res[i] = 0.0 if total == 0.0 else 100.0 * (su - sd) / total
    return res


def mfi(high, low, close, volume, timeperiod=14):  # momentum.rs:286-342
    _cont_slice(high, low, close, volume)
    n = len(high)
    tp = [(high[i] + low[i] + close[i]) / 3.0 for i in range(n)]
    mf = [tp[i] * volume[i] for i in range(n)]
    pos = neg = 0.0
    res = [None] * n
    for i in range(1, n):
        if tp[i] > tp[i - 1]:
            pos += mf[i]
        elif tp[i] < tp[i - 1]:
            neg += mf[i]
        if i >= timeperiod:
            prev = i - timeperiod
            if prev > 0:
                if tp[prev] > tp[prev - 1]:
                    pos -= mf[prev]
                elif tp[prev] < tp[prev - 1]:
                    neg -= mf[prev]
        if i >= timeperiod:
            if neg == 0.0:
                res[i] = 100.0
            else:
                mr = pos / neg
                res[i] = 100.0 - (100.0 / (1.0 + mr))
    return res


def cci(high, low, close, timeperiod=14):  # momentum.rs:138-178
    _cont_slice(high, low, close)
    n = len(high)
    tp = [(high[i] + low[i] + close[i]) / 3.0 for i in range(n)]
    sma_tp = calc_sma(tp, timeperiod)
    res = [None] * n
    if timeperiod == 0:
        return res
    for i in range(timeperiod - 1, n):
        avg = sma_tp[i]
        if avg is None:
            continue
        md = 0.0
        for j in range(i + 1 - timeperiod, i + 1):
            md += abs(tp[j] - avg)
        if md != 0.0:
            md /= float(timeperiod)
            res[i] = (tp[i] - avg) / (0.015 * md)
    return res


# ---------------------------------------------------------------- volatility.rs / volume.rs
def _shift1(col):
    return [None] + list(col[:-1]) if len(col) else []


def calc_trange(high, low, close):  # volatility.rs:67-84
    out = []
    for h, l, pc in zip(high, low, _shift1(close)):
        if h is None or l is None or pc is None:
            out.append(None)
        else:
            out.append(rs_max(rs_max(h - l, abs(h - pc)), abs(l - pc)))
    return out


def atr(high, low, close, timeperiod=14):  # volatility.rs:18-31
    return calc_ema(calc_trange(high, low, close), 2 * timeperiod - 1)


def natr(high, low, close, timeperiod=14):  # volatility.rs:34-48
    a = atr(high, low, close, timeperiod)
    return [None if (x is None or c is None) else (x / c) * 100.0 for x, c in zip(a, close)]


def obv(close, volume):  # volume.rs:70-94
    out, s = [], 0.0
    for pc, c, v in zip(_shift1(close), close, volume):
        if pc is None or c is None or v is None:
            out.append(None)
            continue
        d = pc - c
        if d > 0.0:
            s += v
        elif d < 0.0:
            s -= v
        out.append(s)
    return out


def calc_ad(high, low, close, volume):  # volume.rs:100-126
    out, s = [], 0.0
    for h, l, c, v in zip(high, low, close, volume):
        if h is None or l is None or c is None or v is None:
            out.append(None)
            continue
        diff = h - l
        if diff == 0.0:
            out.append(0.0)
        else:
            s += (2.0 * c - l - h) / diff * v
            out.append(s)
    return out


def adosc(high, low, close, volume, fastperiod=3, slowperiod=10):  # volume.rs:34-67
    ad = calc_ad(high, low, close, volume)
    adl, s = [], 0.0
    for a in ad:
        if a is None:
            adl.append(None)
        else:
            s += a
            adl.append(s)
    f, sl = calc_ema(adl, fastperiod), calc_ema(adl, slowperiod)
    return [None if (x is None or y is None) else x - y for x, y in zip(f, sl)]


# ---------------------------------------------------------------- momentum.py compositions
def _rolling(col, window, is_max):
    """polars Expr.rolling_max/min(window): positional window, min_samples = window."""
    out = []
    for i in range(len(col)):
        if window <= 0 or i + 1 < window:
            out.append(None)
            continue
        w = col[i + 1 - window:i + 1]
        if any(v is None for v in w):
            out.append(None)
        else:
            out.append(max(w) if is_max else min(w))
    return out


def _fastk(high, low, close, k):
    ln, hn = _rolling(low, k, False), _rolling(high, k, True)
    out = []
    for c, lo, hi in zip(close, ln, hn):
        if c is None or lo is None or hi is None:
            out.append(None)
        else:
            num, den = (c - lo) * 100.0, hi - lo
            if den == 0.0:
                out.append(math.nan if (num == 0.0 or math.isnan(num)) else math.copysign(math.inf, num))
            else:
                out.append(num / den)
    return out


def stoch(high, low, close, fastk_period=5, slowk_period=3, slowk_matype=0, slowd_period=3,
          slowd_matype=0):  # momentum.py:178-186
    fk = _fastk(high, low, close, fastk_period)
    slowk = calc_ma(fk, slowk_period, slowk_matype)
    slowd = calc_ma(slowk, slowd_period, slowd_matype)
    return slowk, slowd


def stochf(high, low, close, fastk_period=5, fastd_period=3, fastd_matype=0):  # momentum.py:188-195
    fk = _fastk(high, low, close, fastk_period)
    return fk, calc_ma(fk, fastd_period, fastd_matype)


def kdj(high, low, close, fastk_period=9, k_period=3, d_period=3):  # D3
    k, d = stoch(high, low, close, fastk_period, k_period, 0, d_period, 0)
    j = [None if (a is None or b is None) else 3.0 * a - 2.0 * b for a, b in zip(k, d)]
    return k, d, j
