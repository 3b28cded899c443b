/*
 * pq_oracle.c -- CPU ORACLE for the polars-quant src/talib indicator path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (polars_quant_b200/csrc) never links, loads or falls back to anything in oracle/.
 *
 * What it is: a literal, scalar, single-pass-per-function restatement in plain C of the
 * reference's Rust loops (read-only at /root/reference), one column at a time, with the
 * same multi-pass structure (e.g. MACD = 3 EMA passes + 2 elementwise passes), the same
 * operation order, `fma()` exactly where the Rust uses `mul_add`, plain un-contracted
 * arithmetic elsewhere (build with -ffp-contract=off, no fast-math).
 *
 * PARITY PINNED TO THE REFERENCE'S OWN OUTPUTS (round 2).  The reference ships no tests, golden vectors or
 * KATs for this path (its tests/__init__.py:1-4 is a scratch TA-Lib call) and cannot be compiled here (no Rust
 * toolchain; the snapshot has undefined symbols, e.g. calc_rma).  Its functions are, however, plain scalar
 * loops, so tests/golden/make_ref_golden.py EXECUTES the reference's source text (the .rs files of src/talib through the
 * small Rust-subset interpreter in tests/golden/rustexec/, the .py shims of python/polars_quant/talib imported verbatim)
 * and commits the outputs as tests/golden/talib_ref_golden.npz; tests/test_oracle_ref_golden.py holds every
 * function of this file to those vectors bit for bit (values, validity, and the inputs on which the reference
 * returns Err or panics).  Still defined here rather than by the reference: D1-D3 below (the reference's own
 * gaps).  The older vectors (tests/golden/talib_golden.npz, from the independent restatement oracle/ref_py.py)
 * are kept as a second check.
 *
 * Column convention: `x` = values, `xok` = byte validity (1 = valid, 0 = null) or NULL for
 * "no validity bitmap / no nulls"; outputs `out` / `ok` likewise (ok never NULL).  A null
 * output slot holds NaN in `out`.  All functions return 0, or a negative code when the
 * reference itself would return Err / abort on that input (PQO_ERR_*).
 *
 * Frozen decisions for reference gaps:
 *  D1 calc_rma(x, p): None for i<p-1; out[p-1] = (sum_{j<p} x[j]) / p (left-to-right);
 *     then out[i] = fma(1.0/p, x[i]-out[i-1], out[i-1]).  All None when p==0 || n<p.
 *     (called at momentum.rs:526-527, never defined; mirrors calc_ema overlap.rs:709-723
 *      with the Wilder alpha the reference itself uses at volatility.rs:30.)
 *  D2 slice-style calc_ema/calc_sma(x:&[f64], p) -> Vec<Option<f64>>: the arithmetic of the
 *     no-validity branches overlap.rs:705-724 / :915-931, None where those append null.
 *  D3 KDJ := STOCH(high, low, close, 9, 3, 0, 3, 0) per momentum.py:178-186,
 *     K = slowk, D = slowd, J = 3*K - 2*D (null where K or D null).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PQO_OK 0
#define PQO_ERR_NULLS (-1)     /* reference: cont_slice()? fails on nulls -> PolarsResult::Err */
#define PQO_ERR_PANIC (-2)     /* reference would panic (= abort, Cargo.toml:21 panic="abort") */
#define PQO_ERR_SHAPE (-3)     /* reference: polars arithmetic on unequal lengths -> Err */
#define PQO_ERR_ALLOC (-4)

#define EXPORT __attribute__((visibility("default")))

static inline int is_ok(const uint8_t *ok, int64_t i) { return ok == NULL || ok[i] != 0; }
static inline void put_null(double *out, uint8_t *ok, int64_t i) { out[i] = NAN; ok[i] = 0; }
static inline void put_val(double *out, uint8_t *ok, int64_t i, double v) { out[i] = v; ok[i] = 1; }
static void all_null(double *out, uint8_t *ok, int64_t n) {
    for (int64_t i = 0; i < n; ++i) put_null(out, ok, i);
}
static int has_nulls(const uint8_t *ok, int64_t n) {
    if (!ok) return 0;
    for (int64_t i = 0; i < n; ++i) if (!ok[i]) return 1;
    return 0;
}
/* Rust f64::max / f64::min: if one operand is NaN the other is returned. */
static inline double rs_max(double a, double b) { return fmax(a, b); }
static inline double rs_min(double a, double b) { return fmin(a, b); }

/* ------------------------------------------------------------------ calc_sma
 * overlap.rs:871-937.  count/sum/VecDeque window; nulls skipped (emit null, no state
 * change); out = sum * (1.0/p) (reciprocal multiply, :880,:910). */
EXPORT int pqo_sma(const double *x, const uint8_t *xok, int64_t n, int64_t p,
                   double *out, uint8_t *ok) {
    if (p <= 0 || n < p) { all_null(out, ok, n); return PQO_OK; }     /* :874 */
    double denominator = 1.0 / (double)p;                              /* :880 */
    int64_t count = 0; double sum = 0.0;
    double *window = (double *)malloc(sizeof(double) * (size_t)(p + 1));
    if (!window) return PQO_ERR_ALLOC;
    int64_t head = 0, len = 0;                                          /* ring = VecDeque */
    for (int64_t i = 0; i < n; ++i) {
        if (!is_ok(xok, i)) { put_null(out, ok, i); continue; }        /* :893-896 */
        double value = x[i];
        count += 1; sum += value;                                      /* :898-899 */
        window[(head + len) % (p + 1)] = value; len += 1;              /* push_back */
        if (count < p) { put_null(out, ok, i); }
        else {
            if (count > p) {                                           /* :904-909 */
                double old = window[head]; head = (head + 1) % (p + 1); len -= 1;
                sum -= old; count -= 1;
            }
            put_val(out, ok, i, sum * denominator);                    /* :910 */
        }
    }
    free(window);
    return PQO_OK;
}

/* ------------------------------------------------------------------ calc_ema
 * overlap.rs:660-730.  alpha = 2/(p+1); count<p accumulate; count==p seed = sum/p
 * (emitted); afterwards ema = alpha.mul_add(value - ema, ema). */
EXPORT int pqo_ema(const double *x, const uint8_t *xok, int64_t n, int64_t p,
                   double *out, uint8_t *ok) {
    if (p <= 0 || n < p) { all_null(out, ok, n); return PQO_OK; }     /* :663 */
    double alpha = 2.0 / ((double)p + 1.0);                            /* :669 */
    int64_t count = 0; double ema = 0.0, sum = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (!is_ok(xok, i)) { put_null(out, ok, i); continue; }
        double value = x[i];
        count += 1;
        if (count < p) { sum += value; put_null(out, ok, i); }
        else if (count == p) { sum += value; ema = sum / (double)p; put_val(out, ok, i, ema); }
        else { ema = fma(alpha, value - ema, ema); put_val(out, ok, i, ema); }   /* :698 */
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ D1 calc_rma (frozen) */
EXPORT int pqo_rma(const double *x, int64_t n, int64_t p, double *out, uint8_t *ok) {
    if (p <= 0 || n < p) { all_null(out, ok, n); return PQO_OK; }
    double a = 1.0 / (double)p, sum = 0.0, y = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (i < p - 1) { sum += x[i]; put_null(out, ok, i); }
        else if (i == p - 1) { sum += x[i]; y = sum / (double)p; put_val(out, ok, i, y); }
        else { y = fma(a, x[i] - y, y); put_val(out, ok, i, y); }
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ calc_tema
 * overlap.rs:1177-1311.  Three cascaded EMAs; stage k seeded by the mean of the previous
 * stage's first p outputs; first value at count == 3p-2. */
EXPORT int pqo_tema(const double *x, const uint8_t *xok, int64_t n, int64_t p,
                    double *out, uint8_t *ok) {
    if (p <= 0 || n < 3 * p - 2) { all_null(out, ok, n); return PQO_OK; }   /* :1180 */
    double alpha = 2.0 / ((double)p + 1.0);
    int64_t count = 0; double e[3] = {0, 0, 0}, s[3] = {0, 0, 0};
    for (int64_t i = 0; i < n; ++i) {
        if (!is_ok(xok, i)) { put_null(out, ok, i); continue; }
        double value = x[i];
        count += 1;
        if (count < p) { s[0] += value; put_null(out, ok, i); }
        else if (count == p) {
            s[0] += value; e[0] = s[0] / (double)p; s[1] = e[0]; put_null(out, ok, i);
        } else if (count < 2 * p - 1) {
            e[0] = fma(alpha, value - e[0], e[0]); s[1] += e[0]; put_null(out, ok, i);
        } else if (count == 2 * p - 1) {
            e[0] = fma(alpha, value - e[0], e[0]); s[1] += e[0];
            e[1] = s[1] / (double)p; s[2] = e[1]; put_null(out, ok, i);
        } else if (count < 3 * p - 2) {
            e[0] = fma(alpha, value - e[0], e[0]);
            e[1] = fma(alpha, e[0] - e[1], e[1]);
            s[2] += e[1]; put_null(out, ok, i);
        } else if (count == 3 * p - 2) {
            e[0] = fma(alpha, value - e[0], e[0]);
            e[1] = fma(alpha, e[0] - e[1], e[1]);
            s[2] += e[1]; e[2] = s[2] / (double)p;
            put_val(out, ok, i, 3.0 * e[0] - 3.0 * e[1] + e[2]);       /* :1293 */
        } else {
            e[0] = fma(alpha, value - e[0], e[0]);
            e[1] = fma(alpha, e[0] - e[1], e[1]);
            e[2] = fma(alpha, e[1] - e[2], e[2]);
            put_val(out, ok, i, 3.0 * e[0] - 3.0 * e[1] + e[2]);
        }
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ calc_trima
 * overlap.rs:1313-1326: odd p: n=p/2+1 -> SMA(SMA(x,n),n); even: n=p/2 -> SMA(SMA(x,n),n+1). */
EXPORT int pqo_trima(const double *x, const uint8_t *xok, int64_t n, int64_t p,
                     double *out, uint8_t *ok) {
    double *t = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *tok = (uint8_t *)malloc((size_t)(n + 1));
    if (!t || !tok) { free(t); free(tok); return PQO_ERR_ALLOC; }
    int64_t n1, n2;
    if (p % 2 == 1) { n1 = p / 2 + 1; n2 = n1; } else { n1 = p / 2; n2 = n1 + 1; }
    pqo_sma(x, xok, n, n1, t, tok);
    pqo_sma(t, tok, n, n2, out, ok);
    free(t); free(tok);
    return PQO_OK;
}

/* ------------------------------------------------------------------ calc_wma (oracle-only)
 * overlap.rs:1328-1399, literal (the recurrence is NOT a weighted MA: numerator adds
 * count*value and removes p*old only). */
EXPORT int pqo_wma(const double *x, const uint8_t *xok, int64_t n, int64_t p,
                   double *out, uint8_t *ok) {
    if (p <= 0 || n < p) { all_null(out, ok, n); return PQO_OK; }
    int64_t count = 0;
    double denominator = (double)(p * (p + 1) / 2), numerator = 0.0, sum = 0.0;
    double *window = (double *)malloc(sizeof(double) * (size_t)(p + 1));
    if (!window) return PQO_ERR_ALLOC;
    int64_t head = 0, len = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!is_ok(xok, i)) { put_null(out, ok, i); continue; }
        double value = x[i];
        count += 1; sum += value; numerator += ((double)count) * value;
        window[(head + len) % (p + 1)] = value; len += 1;
        if (count < p) put_null(out, ok, i);
        else {
            if (count > p) {
                double old = window[head]; head = (head + 1) % (p + 1); len -= 1;
                sum -= old; numerator -= ((double)p) * old; count -= 1;
            }
            put_val(out, ok, i, numerator / denominator);
        }
    }
    (void)sum;
    free(window);
    return PQO_OK;
}

/* ------------------------------------------------------------------ calc_dema (oracle-only)
 * overlap.rs:543-658, LITERAL including its defects.  `ema_value` and `sum` are 2-slot ArrayVecs (:554-555).
 * The bitmap branch (:560-602, taken by a chunk that carries a validity bitmap) is a DEMA whose first value
 * appears at count == 2p (one bar late).  The no-bitmap branch (:603-654) is a pasted TEMA body: it indexes
 * slot 2 of the 2-slot arrays (:625 and on) -> panic (= abort) as soon as count reaches 2p-1 (p == 1: at
 * count 2, through the `_` arm).  `xok == NULL` selects the no-bitmap branch, like everywhere in this file. */
EXPORT int pqo_dema(const double *x, const uint8_t *xok, int64_t n, int64_t p,
                    double *out, uint8_t *ok) {
    if (p <= 0 || n < 2 * p - 1) { all_null(out, ok, n); return PQO_OK; }   /* :546 */
    double alpha = 2.0 / ((double)p + 1.0);
    int64_t count = 0; double e[2] = {0, 0}, s[2] = {0, 0};
    for (int64_t i = 0; i < n; ++i) {
        if (!is_ok(xok, i)) { put_null(out, ok, i); continue; }
        double value = x[i];
        count += 1;
        if (count < p) { s[0] += value; put_null(out, ok, i); }
        else if (count == p) { s[0] += value; e[0] = s[0] / (double)p; s[1] = e[0]; put_null(out, ok, i); }
        else if (count < 2 * p - 1) { e[0] = fma(alpha, value - e[0], e[0]); s[1] += e[0]; put_null(out, ok, i); }
        else if (count == 2 * p - 1) {
            e[0] = fma(alpha, value - e[0], e[0]); s[1] += e[0]; e[1] = s[1] / (double)p;
            if (xok == NULL) return PQO_ERR_PANIC;                         /* :625 sum[2] = ... */
            put_null(out, ok, i);
        } else {
            if (xok == NULL) return PQO_ERR_PANIC;                         /* :627-651: slot 2 again */
            e[0] = fma(alpha, value - e[0], e[0]);
            e[1] = fma(alpha, e[0] - e[1], e[1]);
            put_val(out, ok, i, 2.0 * e[0] - e[1]);                        /* :597 */
        }
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ calc_t3 (oracle-only)
 * overlap.rs:939-1175, LITERAL.  Six cascaded EMAs; stage k accumulates while count < (k+1)p-k and is
 * seeded at count == (k+1)p-k -- except the sixth (k = 5), which has no seeding arm (:1042-1057): it starts
 * from 0.0 in the `_` arm.  The two branches end differently: the bitmap branch (:1058-1063) emits
 * 6e0 - 15e1 + 20e2 - 15e3 + 6e4 - e5, the no-bitmap branch (:1160-1166) the nested mul_add of c1..c4 over
 * e5..e2.  The arms are tried in source order (first match wins: matters for p == 1). */
EXPORT int pqo_t3(const double *x, const uint8_t *xok, int64_t n, int64_t p, double vfactor,
                  double *out, uint8_t *ok) {
    if (p <= 0 || n < 6 * p - 5) { all_null(out, ok, n); return PQO_OK; }   /* :942 */
    double alpha = 2.0 / ((double)p + 1.0);
    double v3 = vfactor * vfactor * vfactor, v2 = vfactor * vfactor;          /* powi(3), powi(2) */
    double c1 = -v3;                                                          /* :949 */
    double c2 = 3.0 * v2 - 3.0 * c1;
    double c3 = -2.0 * c2 - 3.0 * c1 - 3.0 * vfactor;
    double c4 = 1.0 - c1 - c2 - c3;
    int64_t count = 0; double e[6] = {0, 0, 0, 0, 0, 0}, s[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < n; ++i) {
        if (!is_ok(xok, i)) { put_null(out, ok, i); continue; }
        double value = x[i];
        count += 1;
        int k, done = 0;
        for (k = 0; k < 6 && !done; ++k) {
            int64_t t = (int64_t)(k + 1) * p - k;
            int acc = count < t, seed = (k < 5) && count == t;
            if (!acc && !seed) continue;
            for (int j = 0; j < k; ++j)                                       /* the stages already running */
                e[j] = fma(alpha, (j == 0 ? value : e[j - 1]) - e[j], e[j]);
            s[k] += (k == 0) ? value : e[k - 1];
            if (seed) { e[k] = s[k] / (double)p; s[k + 1] = e[k]; }
            put_null(out, ok, i);
            done = 1;
        }
        if (done) continue;
        for (int j = 0; j < 6; ++j)
            e[j] = fma(alpha, (j == 0 ? value : e[j - 1]) - e[j], e[j]);
        if (xok != NULL)
            put_val(out, ok, i, 6.0 * e[0] - 15.0 * e[1] + 20.0 * e[2] - 15.0 * e[3] + 6.0 * e[4] - e[5]);
        else
            put_val(out, ok, i, fma(c1, e[5], fma(c2, e[4], fma(c3, e[3], c4 * e[2]))));
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ calc_kama (oracle-only)
 * overlap.rs:732-855, LITERAL.  Pass 1 (:748-807): a non-standard efficiency ratio,
 * er = |v - v[t-p]| / running sum of those same |v - v[t-p]| terms (window_sum), null for the first p values;
 * `window_sum.pop_front().unwrap()` panics for p == 1.  Then sc = (er * (2/3 - 2/31) + 2/31)^2 with polars'
 * null-propagating arithmetic (:811-815).  Pass 2 (:817-852) needs `values.cont_slice().unwrap()`: a column
 * with nulls (or several chunks) panics; the first p non-null sc slots only accumulate the price (null), the
 * next one emits sum/p WITHOUT using its own bar, afterwards kama = sc.mul_add(v - kama, kama). */
EXPORT int pqo_kama(const double *x, const uint8_t *xok, int64_t n, int64_t p,
                    double *out, uint8_t *ok) {
    if (p <= 0 || n < p) { all_null(out, ok, n); return PQO_OK; }           /* :735 */
    double *er = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *erok = (uint8_t *)malloc((size_t)(n + 1));
    double *win = (double *)malloc(sizeof(double) * (size_t)(n + 1));       /* VecDeques as [head, tail) */
    double *wsum = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    if (!er || !erok || !win || !wsum) { free(er); free(erok); free(win); free(wsum); return PQO_ERR_ALLOC; }
    int64_t wh = 0, wt = 0, sh = 0, st = 0, count = 0;
    double sum = 0.0, diff_abs;
    int rc = PQO_OK;
    for (int64_t i = 0; i < n && rc == PQO_OK; ++i) {
        if (!is_ok(xok, i)) { er[i] = NAN; erok[i] = 0; continue; }
        double value = x[i];
        if (count == 0) { count += 1; win[wt++] = value; er[i] = NAN; erok[i] = 0; }
        else if (count < p) {
            count += 1;
            diff_abs = fabs(value - win[wh]);                                  /* front() */
            sum += diff_abs; win[wt++] = value; wsum[st++] = diff_abs;
            er[i] = NAN; erok[i] = 0;
        } else {
            if (wh == wt || sh == st) { rc = PQO_ERR_PANIC; break; }           /* unwrap() on None */
            diff_abs = fabs(value - win[wh++]);
            sum += diff_abs - wsum[sh++];
            win[wt++] = value; wsum[st++] = diff_abs;
            er[i] = diff_abs / sum; erok[i] = 1;
        }
    }
    if (rc == PQO_OK && has_nulls(xok, n)) rc = PQO_ERR_PANIC;                 /* :826 cont_slice().unwrap() */
    if (rc == PQO_OK) {
        double fast_sc = 2.0 / 3.0, slow_sc = 2.0 / 31.0;
        double kama = 0.0; sum = 0.0; count = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (!erok[i]) { put_null(out, ok, i); continue; }
            double sc_sqrt = er[i] * (fast_sc - slow_sc) + slow_sc;
            double sc = sc_sqrt * sc_sqrt;
            if (count < p) { count += 1; sum += x[i]; put_null(out, ok, i); }
            else if (count == p) { count += 1; kama = sum / (double)p; put_val(out, ok, i, kama); }
            else { kama = fma(sc, x[i] - kama, kama); put_val(out, ok, i, kama); }
        }
    }
    free(er); free(erok); free(win); free(wsum);
    return rc;
}

/* ------------------------------------------------------------------ calc_ma
 * overlap.rs:857-869. */
EXPORT int pqo_ma(const double *x, const uint8_t *xok, int64_t n, int64_t p, int64_t matype,
                  double *out, uint8_t *ok) {
    switch (matype) {
        case 1: return pqo_ema(x, xok, n, p, out, ok);
        case 2: return pqo_wma(x, xok, n, p, out, ok);
        case 3: return pqo_dema(x, xok, n, p, out, ok);
        case 4: return pqo_tema(x, xok, n, p, out, ok);
        case 5: return pqo_trima(x, xok, n, p, out, ok);
        case 6: return pqo_kama(x, xok, n, p, out, ok);
        case 8: return pqo_t3(x, xok, n, p, 0.0, out, ok);
        default: return pqo_sma(x, xok, n, p, out, ok);               /* 0, 7, other */
    }
}

/* ------------------------------------------------------------------ bbands
 * overlap.rs:47-116.  mean = sum / p (division); var = sum_sq/p - mean*mean (un-fused);
 * std = sqrt(max(var, 0)). */
EXPORT int pqo_bbands(const double *x, const uint8_t *xok, int64_t n, int64_t p,
                      double nbdevup, double nbdevdn,
                      double *up, uint8_t *upok, double *mid, uint8_t *midok,
                      double *lo, uint8_t *look) {
    if (p <= 0 || n < p) {                                             /* :56-64 */
        all_null(up, upok, n); all_null(mid, midok, n); all_null(lo, look, n);
        return PQO_OK;
    }
    int64_t count = 0; double sum = 0.0, sum_sq = 0.0;
    double *window = (double *)malloc(sizeof(double) * (size_t)(p + 1));
    if (!window) return PQO_ERR_ALLOC;
    int64_t head = 0, len = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!is_ok(xok, i)) {
            put_null(up, upok, i); put_null(mid, midok, i); put_null(lo, look, i); continue;
        }
        double value = x[i];
        count += 1; sum += value; sum_sq += value * value;            /* :84-86 */
        window[(head + len) % (p + 1)] = value; len += 1;
        if (count < p) {
            put_null(up, upok, i); put_null(mid, midok, i); put_null(lo, look, i);
        } else {
            if (count > p) {
                double old = window[head]; head = (head + 1) % (p + 1); len -= 1;
                sum -= old; sum_sq -= old * old; count -= 1;          /* :96-98 */
            }
            double mean = sum / (double)p;                              /* :101 */
            double variance = (sum_sq / (double)p) - mean * mean;       /* :102 */
            double sd = sqrt(rs_max(variance, 0.0));                    /* :103 */
            put_val(up, upok, i, mean + nbdevup * sd);
            put_val(mid, midok, i, mean);
            put_val(lo, look, i, mean - nbdevdn * sd);
        }
    }
    free(window);
    return PQO_OK;
}

/* ------------------------------------------------------------------ monotonic deque helper
 * Literal VecDeque<(usize,f64)> as used by midpoint/midprice (overlap.rs:189-272,292-398). */
typedef struct { uint64_t *idx; double *val; int64_t cap, head, len; } dq_t;
static int dq_init(dq_t *d, int64_t cap) {
    d->cap = cap; d->head = 0; d->len = 0;
    d->idx = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)cap);
    d->val = (double *)malloc(sizeof(double) * (size_t)cap);
    return (d->idx && d->val) ? 0 : -1;
}
static void dq_free(dq_t *d) { free(d->idx); free(d->val); }
static inline int64_t dq_pos(const dq_t *d, int64_t k) { return (d->head + k) % d->cap; }
static void dq_grow(dq_t *d) {
    int64_t ncap = d->cap * 2;
    uint64_t *ni = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)ncap);
    double *nv = (double *)malloc(sizeof(double) * (size_t)ncap);
    for (int64_t k = 0; k < d->len; ++k) { ni[k] = d->idx[dq_pos(d, k)]; nv[k] = d->val[dq_pos(d, k)]; }
    free(d->idx); free(d->val); d->idx = ni; d->val = nv; d->cap = ncap; d->head = 0;
}
static inline void dq_push_back(dq_t *d, uint64_t i, double v) {
    if (d->len == d->cap) dq_grow(d);
    int64_t p = dq_pos(d, d->len); d->idx[p] = i; d->val[p] = v; d->len += 1;
}
static inline void dq_pop_back(dq_t *d) { d->len -= 1; }
static inline void dq_pop_front(dq_t *d) { if (d->len > 0) { d->head = (d->head + 1) % d->cap; d->len -= 1; } }

/* ------------------------------------------------------------------ midpoint
 * overlap.rs:180-278, LITERAL including its defect: the min-side expiry tests
 * window_max.front() (:227,:264), so window_min only expires when the max deque's front
 * index equals count-p *after* the max side already expired it, i.e. never for p>=1:
 * result = (rollmax_p + cummin)/2.  `count - timeperiod` is a wrapping usize (release). */
EXPORT int pqo_midpoint(const double *x, const uint8_t *xok, int64_t n, int64_t p,
                        double *out, uint8_t *ok) {
    dq_t wmax, wmin;
    if (dq_init(&wmax, 64) || dq_init(&wmin, 64)) return PQO_ERR_ALLOC;
    uint64_t count = 0; double mx, mn;
    for (int64_t i = 0; i < n; ++i) {
        if (!is_ok(xok, i)) { put_null(out, ok, i); continue; }
        double value = x[i];
        count += 1;
        while (wmax.len > 0 && wmax.val[dq_pos(&wmax, wmax.len - 1)] <= value) dq_pop_back(&wmax);
        if (wmax.len > 0 && wmax.idx[wmax.head] == count - (uint64_t)p) dq_pop_front(&wmax);
        dq_push_back(&wmax, count, value);
        mx = wmax.val[wmax.head];
        while (wmin.len > 0 && wmin.val[dq_pos(&wmin, wmin.len - 1)] >= value) dq_pop_back(&wmin);
        if (wmax.len > 0 && wmax.idx[wmax.head] == count - (uint64_t)p) dq_pop_front(&wmin);  /* sic */
        dq_push_back(&wmin, count, value);
        mn = wmin.val[wmin.head];
        put_val(out, ok, i, (mx + mn) / 2.0);
    }
    dq_free(&wmax); dq_free(&wmin);
    return PQO_OK;
}

/* ------------------------------------------------------------------ midprice (Donchian mid)
 * overlap.rs:281-404.  Rolling max(high,p) and rolling min(low,p) with expanding start (no
 * warm-up nulls), then (hmax + lmin) / 2.0 via null-propagating polars arithmetic.  The
 * null branch for `low` (:352-376) pushes its nulls into high_builder and pops on `<=`;
 * the two builders then differ in length -> polars arithmetic Err: PQO_ERR_SHAPE. */
EXPORT int pqo_midprice(const double *high, const uint8_t *hok, const double *low,
                        const uint8_t *lok, int64_t n, int64_t p, double *out, uint8_t *ok) {
    if (has_nulls(lok, n)) return PQO_ERR_SHAPE;
    dq_t w;
    if (dq_init(&w, 64)) return PQO_ERR_ALLOC;
    double *hm = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *hmok = (uint8_t *)malloc((size_t)(n + 1));
    if (!hm || !hmok) { free(hm); free(hmok); dq_free(&w); return PQO_ERR_ALLOC; }
    uint64_t count = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!is_ok(hok, i)) { hm[i] = NAN; hmok[i] = 0; continue; }
        double value = high[i];
        count += 1;
        while (w.len > 0 && w.val[dq_pos(&w, w.len - 1)] <= value) dq_pop_back(&w);
        if (w.len > 0 && w.idx[w.head] == count - (uint64_t)p) dq_pop_front(&w);
        dq_push_back(&w, count, value);
        hm[i] = w.val[w.head]; hmok[i] = 1;
    }
    w.head = 0; w.len = 0; count = 0;
    for (int64_t i = 0; i < n; ++i) {
        double value = low[i];
        count += 1;
        /* :363 the bitmap branch of the low pass compares with `<=` (a rolling MAX of low); :384 the no-bitmap
         * branch with `>=`.  A low column that carries a validity bitmap, even an all-set one, takes the former. */
        if (lok) { while (w.len > 0 && w.val[dq_pos(&w, w.len - 1)] <= value) dq_pop_back(&w); }
        else     { while (w.len > 0 && w.val[dq_pos(&w, w.len - 1)] >= value) dq_pop_back(&w); }
        if (w.len > 0 && w.idx[w.head] == count - (uint64_t)p) dq_pop_front(&w);
        dq_push_back(&w, count, value);
        double lm = w.val[w.head];
        if (hmok[i]) put_val(out, ok, i, (hm[i] + lm) / 2.0); else put_null(out, ok, i);   /* :401 */
    }
    free(hm); free(hmok); dq_free(&w);
    return PQO_OK;
}

/* ------------------------------------------------------------------ rsi
 * momentum.rs:507-541 (+ D1).  cont_slice()? -> Err on nulls. */
EXPORT int pqo_rsi(const double *x, const uint8_t *xok, int64_t n, int64_t p,
                   double *out, uint8_t *ok) {
    if (has_nulls(xok, n)) return PQO_ERR_NULLS;
    double *ups = (double *)calloc((size_t)(n + 1), sizeof(double));
    double *downs = (double *)calloc((size_t)(n + 1), sizeof(double));
    double *au = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *ad = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *auok = (uint8_t *)malloc((size_t)(n + 1)), *adok = (uint8_t *)malloc((size_t)(n + 1));
    if (!ups || !downs || !au || !ad || !auok || !adok) {
        free(ups); free(downs); free(au); free(ad); free(auok); free(adok); return PQO_ERR_ALLOC;
    }
    for (int64_t i = 1; i < n; ++i) {                                  /* :515-524 */
        double diff = x[i] - x[i - 1];
        if (diff > 0.0) ups[i] = diff; else downs[i] = -diff;
    }
    pqo_rma(ups, n, p, au, auok);
    pqo_rma(downs, n, p, ad, adok);
    for (int64_t i = 0; i < n; ++i) {                                  /* :529-539 */
        if (auok[i] && adok[i]) {
            if (ad[i] == 0.0) put_val(out, ok, i, 100.0);
            else { double rs = au[i] / ad[i]; put_val(out, ok, i, 100.0 - (100.0 / (1.0 + rs))); }
        } else put_null(out, ok, i);
    }
    free(ups); free(downs); free(au); free(ad); free(auok); free(adok);
    return PQO_OK;
}

/* ------------------------------------------------------------------ macd
 * momentum.rs:250-283 (+ D2).  dea = EMA(dif with None -> 0.0) over the whole array. */
EXPORT int pqo_macd(const double *x, const uint8_t *xok, int64_t n, int64_t fast, int64_t slow,
                    int64_t signal, double *macd, uint8_t *macdok, double *sig, uint8_t *sigok,
                    double *hist, uint8_t *histok) {
    if (has_nulls(xok, n)) return PQO_ERR_NULLS;
    double *f = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *s = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *z = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *fok = (uint8_t *)malloc((size_t)(n + 1)), *sok = (uint8_t *)malloc((size_t)(n + 1));
    if (!f || !s || !z || !fok || !sok) { free(f); free(s); free(z); free(fok); free(sok); return PQO_ERR_ALLOC; }
    pqo_ema(x, NULL, n, fast, f, fok);
    pqo_ema(x, NULL, n, slow, s, sok);
    for (int64_t i = 0; i < n; ++i) {                                  /* :262-266 */
        if (fok[i] && sok[i]) { put_val(macd, macdok, i, f[i] - s[i]); z[i] = macd[i]; }
        else { put_null(macd, macdok, i); z[i] = 0.0; }               /* unwrap_or(0.0) :269 */
    }
    pqo_ema(z, NULL, n, signal, sig, sigok);                           /* :268-271 */
    for (int64_t i = 0; i < n; ++i) {                                  /* :273-277 */
        if (macdok[i] && sigok[i]) put_val(hist, histok, i, macd[i] - sig[i]);
        else put_null(hist, histok, i);
    }
    free(f); free(s); free(z); free(fok); free(sok);
    return PQO_OK;
}

/* ------------------------------------------------------------------ calc_trange
 * volatility.rs:67-84.  pc = close.shift(1) (positional); null unless h, l, pc all valid. */
EXPORT int pqo_trange(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                      const double *c, const uint8_t *cok, int64_t n, double *out, uint8_t *ok) {
    for (int64_t i = 0; i < n; ++i) {
        if (i >= 1 && is_ok(hok, i) && is_ok(lok, i) && is_ok(cok, i - 1)) {
            double pc = c[i - 1];
            double tr = rs_max(rs_max(h[i] - l[i], fabs(h[i] - pc)), fabs(l[i] - pc));   /* :77 */
            put_val(out, ok, i, tr);
        } else put_null(out, ok, i);
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ atr / natr
 * volatility.rs:18-31: calc_ema(trange, 2p-1).  natr :34-48: (atr / close) * 100. */
EXPORT int pqo_atr(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                   const double *c, const uint8_t *cok, int64_t n, int64_t p,
                   double *out, uint8_t *ok) {
    double *tr = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *trok = (uint8_t *)malloc((size_t)(n + 1));
    if (!tr || !trok) { free(tr); free(trok); return PQO_ERR_ALLOC; }
    pqo_trange(h, hok, l, lok, c, cok, n, tr, trok);
    int rc = pqo_ema(tr, trok, n, 2 * p - 1, out, ok);
    free(tr); free(trok);
    return rc;
}
EXPORT int pqo_natr(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                    const double *c, const uint8_t *cok, int64_t n, int64_t p,
                    double *out, uint8_t *ok) {
    int rc = pqo_atr(h, hok, l, lok, c, cok, n, p, out, ok);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        if (ok[i] && is_ok(cok, i)) out[i] = (out[i] / c[i]) * 100.0;  /* :47 */
        else put_null(out, ok, i);
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ obv
 * volume.rs:70-94.  d = close.shift(1) - close (prev - curr); d>0 -> sum += v; d<0 -> sum -= v. */
EXPORT int pqo_obv(const double *c, const uint8_t *cok, const double *v, const uint8_t *vok,
                   int64_t n, double *out, uint8_t *ok) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (i >= 1 && is_ok(cok, i - 1) && is_ok(cok, i) && is_ok(vok, i)) {
            double d = c[i - 1] - c[i];
            if (d > 0.0) sum += v[i]; else if (d < 0.0) sum -= v[i];
            put_val(out, ok, i, sum);
        } else put_null(out, ok, i);
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ calc_ad / ad
 * volume.rs:100-126.  diff==0 -> emit literal 0.0 (sum unchanged). */
EXPORT int pqo_ad(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                  const double *c, const uint8_t *cok, const double *v, const uint8_t *vok,
                  int64_t n, double *out, uint8_t *ok) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (is_ok(hok, i) && is_ok(lok, i) && is_ok(cok, i) && is_ok(vok, i)) {
            double diff = h[i] - l[i];
            if (diff == 0.0) put_val(out, ok, i, 0.0);
            else { sum += (2.0 * c[i] - l[i] - h[i]) / diff * v[i]; put_val(out, ok, i, sum); }
        } else put_null(out, ok, i);
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ adosc
 * volume.rs:34-67: adl = cumsum(calc_ad) (a second cumulative sum), EMA(adl,fast)-EMA(adl,slow). */
EXPORT int pqo_adosc(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                     const double *c, const uint8_t *cok, const double *v, const uint8_t *vok,
                     int64_t n, int64_t fast, int64_t slow, double *out, uint8_t *ok) {
    double *adl = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *ef = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *adlok = (uint8_t *)malloc((size_t)(n + 1)), *efok = (uint8_t *)malloc((size_t)(n + 1));
    if (!adl || !ef || !adlok || !efok) { free(adl); free(ef); free(adlok); free(efok); return PQO_ERR_ALLOC; }
    pqo_ad(h, hok, l, lok, c, cok, v, vok, n, adl, adlok);
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) if (adlok[i]) { sum += adl[i]; adl[i] = sum; }   /* :47-59 */
    pqo_ema(adl, adlok, n, fast, ef, efok);
    pqo_ema(adl, adlok, n, slow, out, ok);
    for (int64_t i = 0; i < n; ++i) {
        if (efok[i] && ok[i]) out[i] = ef[i] - out[i]; else put_null(out, ok, i);    /* :65 */
    }
    free(adl); free(ef); free(adlok); free(efok);
    return PQO_OK;
}

/* ------------------------------------------------------------------ calc_dm and the DM / DI / DX / ADX family
 * momentum.rs:668-727 (calc_dm), :11-61 (adx, adxr), :226-237 (dx), :344-436 (minus_di, minus_dm, plus_di,
 * plus_dm) + D1 (calc_rma).  p_dm[0] = m_dm[0] = tr[0] = 0; the three Wilder averages; DI = 100 * S / ST, null
 * where ST == 0; dx = 100 |p - m| / (p + m) (0 where the sum is 0).  calc_dm returns (dx, minus_di): the
 * reference's `plus_di` takes `.0`, i.e. it returns DX (:409) -- restated literally: out_plus_di == dx.
 * adx = calc_rma(dx with None -> 0.0); adxr[i] = (adx[i] + adx[i - (p-1)]) * 0.5 from i = p-1.
 * Nulls in any input: cont_slice()? -> error. */
EXPORT int pqo_dm(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok, const double *c,
                  const uint8_t *cok, int64_t n, int64_t p,
                  double *plus_dm, uint8_t *plus_dm_ok, double *minus_dm, uint8_t *minus_dm_ok,
                  double *dx, uint8_t *dx_ok, double *minus_di, uint8_t *minus_di_ok,
                  double *adx, uint8_t *adx_ok, double *adxr, uint8_t *adxr_ok) {
    if (has_nulls(hok, n) || has_nulls(lok, n) || has_nulls(cok, n)) return PQO_ERR_NULLS;
    double *pd = (double *)calloc((size_t)(n + 1), sizeof(double)), *md = (double *)calloc((size_t)(n + 1), sizeof(double));
    double *tr = (double *)calloc((size_t)(n + 1), sizeof(double)), *st = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *dx0 = (double *)calloc((size_t)(n + 1), sizeof(double));
    uint8_t *stok = (uint8_t *)malloc((size_t)(n + 1));
    if (!pd || !md || !tr || !st || !dx0 || !stok) { free(pd); free(md); free(tr); free(st); free(dx0); free(stok); return PQO_ERR_ALLOC; }
    for (int64_t i = 1; i < n; ++i) {                                                   /* :680-699 */
        double up_move = h[i] - h[i - 1], down_move = l[i - 1] - l[i];
        if (up_move > down_move && up_move > 0.0) pd[i] = up_move;
        if (down_move > up_move && down_move > 0.0) md[i] = down_move;
        tr[i] = rs_max(rs_max(h[i] - l[i], fabs(h[i] - c[i - 1])), fabs(l[i] - c[i - 1]));
    }
    pqo_rma(pd, n, p, plus_dm, plus_dm_ok);                                            /* :701-703; also plus_dm :418-435 */
    pqo_rma(md, n, p, minus_dm, minus_dm_ok);
    pqo_rma(tr, n, p, st, stok);
    for (int64_t i = 0; i < n; ++i) {                                                   /* :705-725 */
        if (plus_dm_ok[i] && minus_dm_ok[i] && stok[i] && st[i] != 0.0) {
            double p_di = 100.0 * plus_dm[i] / st[i], m_di = 100.0 * minus_dm[i] / st[i];
            put_val(minus_di, minus_di_ok, i, m_di);
            double diff = fabs(p_di - m_di), sum = p_di + m_di;
            double d = (sum == 0.0) ? 0.0 : 100.0 * diff / sum;
            put_val(dx, dx_ok, i, d);
            dx0[i] = d;
        } else { put_null(minus_di, minus_di_ok, i); put_null(dx, dx_ok, i); }
    }
    pqo_rma(dx0, n, p, adx, adx_ok);                                                    /* :20-26 */
    for (int64_t i = 0; i < n; ++i) put_null(adxr, adxr_ok, i);
    if (p >= 1)
        for (int64_t i = p - 1; i < n; ++i) {                                           /* :49-57 */
            int64_t j = i - (p - 1);                                                    /* saturating_sub: i >= p-1 here */
            if (adx_ok[i] && adx_ok[j]) put_val(adxr, adxr_ok, i, (adx[i] + adx[j]) * 0.5);
        }
    free(pd); free(md); free(tr); free(st); free(dx0); free(stok);
    return PQO_OK;
}

/* ------------------------------------------------------------------ trix
 * momentum.rs:544-571 (+ D2 slice-style calc_ema).  ema2 / ema3 run over the WHOLE previous array with
 * None -> 0.0, so all three seed at index p-1; res[i] = (e3[i] - e3[i-1]) / e3[i-1] * 100, null where e3[i-1] == 0. */
EXPORT int pqo_trix(const double *x, const uint8_t *xok, int64_t n, int64_t p, double *out, uint8_t *ok) {
    if (has_nulls(xok, n)) return PQO_ERR_NULLS;
    double *e = (double *)malloc(sizeof(double) * (size_t)(3 * n + 3)), *z = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *k = (uint8_t *)malloc((size_t)(3 * n + 3));
    if (!e || !z || !k) { free(e); free(z); free(k); return PQO_ERR_ALLOC; }
    double *e1 = e, *e2 = e + n, *e3 = e + 2 * n; uint8_t *k1 = k, *k2 = k + n, *k3 = k + 2 * n;
    pqo_ema(x, NULL, n, p, e1, k1);
    for (int64_t i = 0; i < n; ++i) z[i] = k1[i] ? e1[i] : 0.0;
    pqo_ema(z, NULL, n, p, e2, k2);
    for (int64_t i = 0; i < n; ++i) z[i] = k2[i] ? e2[i] : 0.0;
    pqo_ema(z, NULL, n, p, e3, k3);
    for (int64_t i = 0; i < n; ++i) put_null(out, ok, i);
    for (int64_t i = 1; i < n; ++i)
        if (k3[i] && k3[i - 1] && e3[i - 1] != 0.0) put_val(out, ok, i, (e3[i] - e3[i - 1]) / e3[i - 1] * 100.0);
    free(e); free(z); free(k);
    return PQO_OK;
}

/* ------------------------------------------------------------------ ultosc
 * momentum.rs:573-627.  bp[0] = tr[0] = 0; running sums over p1 / p2 / p3 bars (from index 0); an average is
 * null where its range sum is 0; 100 * (4 a1 + 2 a2 + a3) / 7.  Periods must be >= 1 (usize underflow otherwise). */
EXPORT int pqo_ultosc(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok, const double *c,
                      const uint8_t *cok, int64_t n, int64_t p1, int64_t p2, int64_t p3, double *out, uint8_t *ok) {
    if (has_nulls(hok, n) || has_nulls(lok, n) || has_nulls(cok, n)) return PQO_ERR_NULLS;
    if (p1 < 1 || p2 < 1 || p3 < 1) return PQO_ERR_SHAPE;
    double *bp = (double *)calloc((size_t)(n + 1), sizeof(double)), *tr = (double *)calloc((size_t)(n + 1), sizeof(double));
    double *a = (double *)malloc(sizeof(double) * (size_t)(3 * n + 3));
    uint8_t *ak = (uint8_t *)calloc((size_t)(3 * n + 3), 1);
    if (!bp || !tr || !a || !ak) { free(bp); free(tr); free(a); free(ak); return PQO_ERR_ALLOC; }
    for (int64_t i = 1; i < n; ++i) {
        double min_l_pc = fmin(l[i], c[i - 1]), max_h_pc = fmax(h[i], c[i - 1]);        /* Rust f64::min / max */
        bp[i] = c[i] - min_l_pc;
        tr[i] = max_h_pc - min_l_pc;
    }
    const int64_t per[3] = {p1, p2, p3};
    for (int q = 0; q < 3; ++q) {                                                       /* fn avg :598-613 */
        double s_bp = 0.0, s_tr = 0.0; const int64_t p = per[q];
        for (int64_t i = 0; i < n; ++i) {
            s_bp += bp[i]; s_tr += tr[i];
            if (i >= p) { s_bp -= bp[i - p]; s_tr -= tr[i - p]; }
            if (i >= p - 1 && s_tr != 0.0) { a[q * n + i] = s_bp / s_tr; ak[q * n + i] = 1; }
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        if (ak[i] && ak[n + i] && ak[2 * n + i]) put_val(out, ok, i, 100.0 * (4.0 * a[i] + 2.0 * a[n + i] + a[2 * n + i]) / 7.0);
        else put_null(out, ok, i);
    }
    free(bp); free(tr); free(a); free(ak);
    return PQO_OK;
}

/* ------------------------------------------------------------------ aroon
 * momentum.rs:63-110.  For i >= p: last position of the maximum of high / minimum of low (`>=` / `<=` scans from
 * f64::MIN / f64::MAX) in the p+1 bars [i-p, i], divided by p, times 100.  p must be >= 1. */
EXPORT int pqo_aroon(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok, int64_t n, int64_t p,
                     double *up, uint8_t *upok, double *down, uint8_t *downok) {
    if (has_nulls(hok, n) || has_nulls(lok, n)) return PQO_ERR_NULLS;
    if (p < 1) return PQO_ERR_SHAPE;
    for (int64_t i = 0; i < n; ++i) { put_null(up, upok, i); put_null(down, downok, i); }
    for (int64_t i = p; i < n; ++i) {
        int64_t start = i - p, max_idx = 0, min_idx = 0;
        double max_val = -DBL_MAX, min_val = DBL_MAX;
        for (int64_t j = start; j <= i; ++j) {
            if (h[j] >= max_val) { max_val = h[j]; max_idx = j - start; }
            if (l[j] <= min_val) { min_val = l[j]; min_idx = j - start; }
        }
        put_val(up, upok, i, ((double)max_idx / (double)p) * 100.0);
        put_val(down, downok, i, ((double)min_idx / (double)p) * 100.0);
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ info (last-row reductions)
 * README.md:832-851 `Selector.info()`: README-only in the reference (no source), so this is the DEFINITION the GPU
 * kernel (csrc/info_host.inc) is checked against: T = n-1, `start` = first valid row, sums oldest -> newest.
 * out[13]: price, high, low, volume, return_1d, return_5d, return_20d, volatility, ma_5, ma_10, ma_20,
 * volume_ratio, amplitude; ok[k] = 0 (and NaN) where the statistic needs rows before `start`. */
EXPORT int pqo_info(const double *c, const double *h, const double *l, const double *v, int64_t n, int64_t start,
                    double *out, uint8_t *ok) {
    if (n < 1 || start < 0) return PQO_ERR_SHAPE;
    const int64_t T = n - 1, have = n - start;
    static const int64_t lags[3] = {1, 5, 20}, mas[3] = {5, 10, 20};
    for (int k = 0; k < 13; ++k) put_null(out, ok, k);
    if (have < 1) return PQO_OK;
    put_val(out, ok, 0, c[T]); put_val(out, ok, 1, h[T]); put_val(out, ok, 2, l[T]); put_val(out, ok, 3, v[T]);
    for (int i = 0; i < 3; ++i)
        if (have > lags[i]) put_val(out, ok, 4 + i, (c[T] / c[T - lags[i]] - 1.0) * 100.0);
    if (have >= 21) {
        double r[20], sum = 0.0, ss = 0.0;
        for (int i = 0; i < 20; ++i) { r[i] = c[T - 19 + i] / c[T - 20 + i] - 1.0; sum += r[i]; }
        const double m = sum / 20.0;
        for (int i = 0; i < 20; ++i) { const double d = r[i] - m; ss += d * d; }
        put_val(out, ok, 7, sqrt(ss / 19.0) * sqrt(252.0) * 100.0);
    }
    for (int i = 0; i < 3; ++i)
        if (have >= mas[i]) {
            double sum = 0.0;
            for (int64_t t = T - mas[i] + 1; t <= T; ++t) sum += c[t];
            put_val(out, ok, 8 + i, sum / (double)mas[i]);
        }
    if (have >= 5) {
        double sum = 0.0;
        for (int64_t t = T - 4; t <= T; ++t) sum += v[t];
        put_val(out, ok, 11, v[T] / (sum / 5.0));
    }
    put_val(out, ok, 12, (h[T] - l[T]) / c[T] * 100.0);
    return PQO_OK;
}

/* ------------------------------------------------------------------ willr
 * momentum.rs:630-662.  Brute-force window, f64::MIN/MAX init, Rust max/min. */
EXPORT int pqo_willr(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                     const double *c, const uint8_t *cok, int64_t n, int64_t p,
                     double *out, uint8_t *ok) {
    if (has_nulls(hok, n) || has_nulls(lok, n) || has_nulls(cok, n)) return PQO_ERR_NULLS;
    all_null(out, ok, n);
    if (p <= 0) return PQO_OK;            /* (0usize - 1) wraps: empty range in release */
    for (int64_t i = p - 1; i < n; ++i) {
        double max_h = -1.7976931348623157e308, min_l = 1.7976931348623157e308;
        for (int64_t j = i + 1 - p; j <= i; ++j) { max_h = rs_max(max_h, h[j]); min_l = rs_min(min_l, l[j]); }
        double diff = max_h - min_l;
        put_val(out, ok, i, diff == 0.0 ? 0.0 : -100.0 * (max_h - c[i]) / diff);     /* :653-657 */
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ polars rolling_min/max
 * Third-party: polars 1.39.3 Expr.rolling_min/rolling_max(window_size) (uv.lock:213-214),
 * absent from /root/reference; restated from its published semantics: positional window of
 * `window` rows, min_samples = window, so the result is null unless all `window` rows are
 * non-null (first window-1 rows null).  Call sites: momentum.py:181-182,191-192,201-202. */
static void rolling_ext(const double *x, const uint8_t *xok, int64_t n, int64_t w, int is_max,
                        double *out, uint8_t *ok) {
    for (int64_t i = 0; i < n; ++i) {
        if (w <= 0 || i + 1 < w) { put_null(out, ok, i); continue; }
        int good = 1; double m = x[i + 1 - w];
        for (int64_t j = i + 1 - w; j <= i; ++j) {
            if (!is_ok(xok, j)) { good = 0; break; }
            if (is_max ? (x[j] > m) : (x[j] < m)) m = x[j];
        }
        if (good) put_val(out, ok, i, m); else put_null(out, ok, i);
    }
}

/* fastk = (close - ln) * 100.0 / (hn - ln)  (momentum.py:183,193): null-propagating polars
 * float arithmetic; x/0 is IEEE (+-inf / NaN), not guarded. */
static int fastk_line(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                      const double *c, const uint8_t *cok, int64_t n, int64_t k,
                      double *fk, uint8_t *fkok) {
    double *hn = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *ln = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *hnok = (uint8_t *)malloc((size_t)(n + 1)), *lnok = (uint8_t *)malloc((size_t)(n + 1));
    if (!hn || !ln || !hnok || !lnok) { free(hn); free(ln); free(hnok); free(lnok); return PQO_ERR_ALLOC; }
    rolling_ext(l, lok, n, k, 0, ln, lnok);
    rolling_ext(h, hok, n, k, 1, hn, hnok);
    for (int64_t i = 0; i < n; ++i) {
        if (is_ok(cok, i) && lnok[i] && hnok[i]) put_val(fk, fkok, i, (c[i] - ln[i]) * 100.0 / (hn[i] - ln[i]));
        else put_null(fk, fkok, i);
    }
    free(hn); free(ln); free(hnok); free(lnok);
    return PQO_OK;
}

/* ------------------------------------------------------------------ STOCH / STOCHF / KDJ
 * momentum.py:178-186 / :188-195; KDJ per D3. */
EXPORT int pqo_stoch(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                     const double *c, const uint8_t *cok, int64_t n, int64_t fastk_period,
                     int64_t slowk_period, int64_t slowk_matype, int64_t slowd_period,
                     int64_t slowd_matype, double *slowk, uint8_t *slowkok,
                     double *slowd, uint8_t *slowdok) {
    double *fk = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *fkok = (uint8_t *)malloc((size_t)(n + 1));
    if (!fk || !fkok) { free(fk); free(fkok); return PQO_ERR_ALLOC; }
    int rc = fastk_line(h, hok, l, lok, c, cok, n, fastk_period, fk, fkok);
    if (!rc) rc = pqo_ma(fk, fkok, n, slowk_period, slowk_matype, slowk, slowkok);
    if (!rc) rc = pqo_ma(slowk, slowkok, n, slowd_period, slowd_matype, slowd, slowdok);
    free(fk); free(fkok);
    return rc;
}
EXPORT int pqo_stochf(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                      const double *c, const uint8_t *cok, int64_t n, int64_t fastk_period,
                      int64_t fastd_period, int64_t fastd_matype, double *fastk, uint8_t *fastkok,
                      double *fastd, uint8_t *fastdok) {
    int rc = fastk_line(h, hok, l, lok, c, cok, n, fastk_period, fastk, fastkok);
    if (!rc) rc = pqo_ma(fastk, fastkok, n, fastd_period, fastd_matype, fastd, fastdok);
    return rc;
}
/* STOCHRSI (momentum.py:197-205): the STOCHF construction over the RSI line itself:
 * rsi = RSI(real, timeperiod); fastk = (rsi - rolling_min(rsi, k)) * 100 / (rolling_max(rsi, k) - rolling_min(rsi, k));
 * fastd = MA(fastk, fastd_period, fastd_matype). */
EXPORT int pqo_rsi(const double *x, const uint8_t *xok, int64_t n, int64_t p, double *out, uint8_t *ok);
EXPORT int pqo_stochrsi(const double *x, const uint8_t *xok, int64_t n, int64_t timeperiod, int64_t fastk_period,
                        int64_t fastd_period, int64_t fastd_matype, double *fastk, uint8_t *fastkok,
                        double *fastd, uint8_t *fastdok) {
    double *r = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *rok = (uint8_t *)malloc((size_t)(n + 1));
    if (!r || !rok) { free(r); free(rok); return PQO_ERR_ALLOC; }
    int rc = pqo_rsi(x, xok, n, timeperiod, r, rok);
    if (!rc) rc = fastk_line(r, rok, r, rok, r, rok, n, fastk_period, fastk, fastkok);
    if (!rc) rc = pqo_ma(fastk, fastkok, n, fastd_period, fastd_matype, fastd, fastdok);
    free(r); free(rok);
    return rc;
}
/* MACDEXT (momentum.py:83-88): macd = MA(real, fast, fastmatype) - MA(real, slow, slowmatype) (null where
 * either is null), signal = MA(macd, signalperiod, signalmatype) (null-skipping, as every calc_ma),
 * hist = macd - signal. */
EXPORT int pqo_macdext(const double *x, const uint8_t *xok, int64_t n, int64_t fast, int64_t fastmatype,
                       int64_t slow, int64_t slowmatype, int64_t signal, int64_t signalmatype,
                       double *macd, uint8_t *macdok, double *sig, uint8_t *sigok, double *hist, uint8_t *histok) {
    double *a = (double *)malloc(sizeof(double) * (size_t)(n + 1)), *b = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *aok = (uint8_t *)malloc((size_t)(n + 1)), *bok = (uint8_t *)malloc((size_t)(n + 1));
    if (!a || !b || !aok || !bok) { free(a); free(b); free(aok); free(bok); return PQO_ERR_ALLOC; }
    int rc = pqo_ma(x, xok, n, fast, fastmatype, a, aok);
    if (!rc) rc = pqo_ma(x, xok, n, slow, slowmatype, b, bok);
    if (!rc) {
        for (int64_t i = 0; i < n; ++i) {
            if (aok[i] && bok[i]) put_val(macd, macdok, i, a[i] - b[i]); else put_null(macd, macdok, i);
        }
        rc = pqo_ma(macd, macdok, n, signal, signalmatype, sig, sigok);
    }
    if (!rc) {
        for (int64_t i = 0; i < n; ++i) {
            if (macdok[i] && sigok[i]) put_val(hist, histok, i, macd[i] - sig[i]); else put_null(hist, histok, i);
        }
    }
    free(a); free(b); free(aok); free(bok);
    return rc;
}
EXPORT int pqo_kdj(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                   const double *c, const uint8_t *cok, int64_t n, int64_t fastk_period,
                   int64_t k_period, int64_t d_period, double *K, uint8_t *Kok,
                   double *D, uint8_t *Dok, double *J, uint8_t *Jok) {
    int rc = pqo_stoch(h, hok, l, lok, c, cok, n, fastk_period, k_period, 0, d_period, 0, K, Kok, D, Dok);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        if (Kok[i] && Dok[i]) put_val(J, Jok, i, 3.0 * K[i] - 2.0 * D[i]); else put_null(J, Jok, i);
    }
    return PQO_OK;
}

/* Donchian channel (D3): upper = rollmax_p(high), lower = rollmin_p(low) with the expanding
 * start of midprice (overlap.rs:325-345,378-398); mid = pqo_midprice. */
EXPORT int pqo_donchian(const double *high, const double *low, int64_t n, int64_t p,
                        double *upper, uint8_t *upok, double *lower, uint8_t *look) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t j0 = (p > 0 && i + 1 - p > 0) ? i + 1 - p : 0;
        double mx = high[j0], mn = low[j0];
        for (int64_t j = j0; j <= i; ++j) { if (high[j] > mx) mx = high[j]; if (low[j] < mn) mn = low[j]; }
        put_val(upper, upok, i, mx); put_val(lower, look, i, mn);
    }
    return PQO_OK;
}

/* ------------------------------------------------------------------ riders
 * mom momentum.rs:384-397; roc/rocp/rocr/rocr100 :439-504 (null when prev == 0). */
EXPORT int pqo_mom(const double *x, const uint8_t *xok, int64_t n, int64_t p, double *out, uint8_t *ok) {
    if (has_nulls(xok, n)) return PQO_ERR_NULLS;
    all_null(out, ok, n);
    if (p < 0) return PQO_OK;
    for (int64_t i = p; i < n; ++i) put_val(out, ok, i, x[i] - x[i - p]);
    return PQO_OK;
}
/* kind: 0 roc, 1 rocp, 2 rocr, 3 rocr100 */
EXPORT int pqo_roc(const double *x, const uint8_t *xok, int64_t n, int64_t p, int kind,
                   double *out, uint8_t *ok) {
    if (has_nulls(xok, n)) return PQO_ERR_NULLS;
    all_null(out, ok, n);
    if (p < 0) return PQO_OK;
    for (int64_t i = p; i < n; ++i) {
        double curr = x[i], prev = x[i - p];
        if (prev != 0.0) {
            double r;
            switch (kind) {
                case 0: r = (curr - prev) / prev * 100.0; break;
                case 1: r = (curr - prev) / prev; break;
                case 2: r = curr / prev; break;
                default: r = (curr / prev) * 100.0; break;
            }
            put_val(out, ok, i, r);
        }
    }
    return PQO_OK;
}

/* cmo momentum.rs:181-223. */
EXPORT int pqo_cmo(const double *x, const uint8_t *xok, int64_t n, int64_t p, double *out, uint8_t *ok) {
    if (has_nulls(xok, n)) return PQO_ERR_NULLS;
    double *ups = (double *)calloc((size_t)(n + 1), sizeof(double));
    double *downs = (double *)calloc((size_t)(n + 1), sizeof(double));
    if (!ups || !downs) { free(ups); free(downs); return PQO_ERR_ALLOC; }
    for (int64_t i = 1; i < n; ++i) {
        double diff = x[i] - x[i - 1];
        if (diff > 0.0) ups[i] = diff; else downs[i] = -diff;
    }
    all_null(out, ok, n);
    double su = 0.0, sd = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        su += ups[i]; sd += downs[i];
        if (i >= p) { su -= ups[i - p]; sd -= downs[i - p]; }
        if (p >= 1 && i >= p - 1) {
            double total = su + sd;
            put_val(out, ok, i, total == 0.0 ? 0.0 : 100.0 * (su - sd) / total);
        }
    }
    free(ups); free(downs);
    return PQO_OK;
}

/* mfi momentum.rs:286-342. */
EXPORT int pqo_mfi(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                   const double *c, const uint8_t *cok, const double *v, const uint8_t *vok,
                   int64_t n, int64_t p, double *out, uint8_t *ok) {
    if (has_nulls(hok, n) || has_nulls(lok, n) || has_nulls(cok, n) || has_nulls(vok, n)) return PQO_ERR_NULLS;
    double *tp = (double *)calloc((size_t)(n + 1), sizeof(double));
    double *mf = (double *)calloc((size_t)(n + 1), sizeof(double));
    if (!tp || !mf) { free(tp); free(mf); return PQO_ERR_ALLOC; }
    for (int64_t i = 0; i < n; ++i) { tp[i] = (h[i] + l[i] + c[i]) / 3.0; mf[i] = tp[i] * v[i]; }
    all_null(out, ok, n);
    double pos = 0.0, neg = 0.0;
    for (int64_t i = 1; i < n; ++i) {
        if (tp[i] > tp[i - 1]) pos += mf[i]; else if (tp[i] < tp[i - 1]) neg += mf[i];
        if (i >= p) {
            int64_t prev = i - p;
            if (prev > 0) {
                if (tp[prev] > tp[prev - 1]) pos -= mf[prev]; else if (tp[prev] < tp[prev - 1]) neg -= mf[prev];
            }
            if (neg == 0.0) put_val(out, ok, i, 100.0);
            else { double mr = pos / neg; put_val(out, ok, i, 100.0 - (100.0 / (1.0 + mr))); }
        }
    }
    free(tp); free(mf);
    return PQO_OK;
}

/* cci momentum.rs:138-178 (+ D2 slice-style calc_sma). */
EXPORT int pqo_cci(const double *h, const uint8_t *hok, const double *l, const uint8_t *lok,
                   const double *c, const uint8_t *cok, int64_t n, int64_t p, double *out, uint8_t *ok) {
    if (has_nulls(hok, n) || has_nulls(lok, n) || has_nulls(cok, n)) return PQO_ERR_NULLS;
    double *tp = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *sm = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    uint8_t *smok = (uint8_t *)malloc((size_t)(n + 1));
    if (!tp || !sm || !smok) { free(tp); free(sm); free(smok); return PQO_ERR_ALLOC; }
    for (int64_t i = 0; i < n; ++i) tp[i] = (h[i] + l[i] + c[i]) / 3.0;
    pqo_sma(tp, NULL, n, p, sm, smok);
    all_null(out, ok, n);
    if (p >= 1) for (int64_t i = p - 1; i < n; ++i) {
        if (!smok[i]) continue;
        double avg = sm[i], md = 0.0;
        for (int64_t j = i + 1 - p; j <= i; ++j) md += fabs(tp[j] - avg);
        if (md != 0.0) { md /= (double)p; put_val(out, ok, i, (tp[i] - avg) / (0.015 * md)); }
    }
    free(tp); free(sm); free(smok);
    return PQO_OK;
}

/* ------------------------------------------------------------------ the 15-indicator suite
 * Runs the reference's per-column functions one after another over one symbol, exactly as
 * 15 separate plugin calls would (SURVEY 8a "Proposed 15-indicator fused suite").
 * Output order (21 columns) == include/pqb200.h enum pqb_output. */
typedef struct {
    int32_t sma, ema, tema, trima, bb; double bb_up, bb_dn;
    int32_t macd_fast, macd_slow, macd_signal, rsi, atr, natr;
    int32_t stoch_k, stoch_sk, stoch_sd, willr, midprice;
} pqo_suite_params;

EXPORT int pqo_suite(const double *c, const double *h, const double *l, const double *v, int64_t n,
                     const pqo_suite_params *P, double *const *out, uint8_t *const *ok) {
    int rc = 0;
    rc |= pqo_sma(c, NULL, n, P->sma, out[0], ok[0]);
    rc |= pqo_ema(c, NULL, n, P->ema, out[1], ok[1]);
    rc |= pqo_tema(c, NULL, n, P->tema, out[2], ok[2]);
    rc |= pqo_trima(c, NULL, n, P->trima, out[3], ok[3]);
    rc |= pqo_bbands(c, NULL, n, P->bb, P->bb_up, P->bb_dn, out[4], ok[4], out[5], ok[5], out[6], ok[6]);
    rc |= pqo_macd(c, NULL, n, P->macd_fast, P->macd_slow, P->macd_signal, out[7], ok[7], out[8], ok[8], out[9], ok[9]);
    rc |= pqo_rsi(c, NULL, n, P->rsi, out[10], ok[10]);
    rc |= pqo_trange(h, NULL, l, NULL, c, NULL, n, out[11], ok[11]);
    rc |= pqo_atr(h, NULL, l, NULL, c, NULL, n, P->atr, out[12], ok[12]);
    rc |= pqo_natr(h, NULL, l, NULL, c, NULL, n, P->natr, out[13], ok[13]);
    rc |= pqo_obv(c, NULL, v, NULL, n, out[14], ok[14]);
    rc |= pqo_ad(h, NULL, l, NULL, c, NULL, v, NULL, n, out[15], ok[15]);
    rc |= pqo_kdj(h, NULL, l, NULL, c, NULL, n, P->stoch_k, P->stoch_sk, P->stoch_sd,
                  out[16], ok[16], out[17], ok[17], out[18], ok[18]);
    rc |= pqo_willr(h, NULL, l, NULL, c, NULL, n, P->willr, out[19], ok[19]);
    rc |= pqo_midprice(h, NULL, l, NULL, n, P->midprice, out[20], ok[20]);
    return rc;
}

/* Whole-panel driver for the CPU baseline: panel laid out [field][symbol][pitch] like the GPU
 * panel; outputs [output][symbol][pitch]; pthreads over symbols (one "plugin call chain" per
 * column, as polars' rayon pool would run it).  Returns the number of threads used (>0) or a
 * negative error. */
#include <pthread.h>
#include <unistd.h>
typedef struct {
    const double *c, *h, *l, *v; int64_t n_symbols, n_bars, pitch;
    const pqo_suite_params *P; double *out; uint8_t *ok;
    int64_t *next; int err;
} pqo_job;
static void *pqo_worker(void *arg) {
    pqo_job *J = (pqo_job *)arg;
    const int64_t plane = J->n_symbols * J->pitch;
    for (;;) {
        int64_t s0 = __atomic_fetch_add(J->next, 8, __ATOMIC_RELAXED);   /* dynamic, chunk 8 */
        if (s0 >= J->n_symbols) break;
        int64_t s1 = s0 + 8 < J->n_symbols ? s0 + 8 : J->n_symbols;
        for (int64_t s = s0; s < s1; ++s) {
            double *o[21]; uint8_t *k[21];
            for (int j = 0; j < 21; ++j) { o[j] = J->out + j * plane + s * J->pitch; k[j] = J->ok + j * plane + s * J->pitch; }
            J->err |= pqo_suite(J->c + s * J->pitch, J->h + s * J->pitch, J->l + s * J->pitch,
                                J->v + s * J->pitch, J->n_bars, J->P, o, k);
        }
    }
    return NULL;
}
EXPORT int pqo_suite_panel(const double *c, const double *h, const double *l, const double *v,
                           int64_t n_symbols, int64_t n_bars, int64_t pitch,
                           const pqo_suite_params *P, double *out, uint8_t *ok, int threads) {
    if (threads <= 0) { long nc = sysconf(_SC_NPROCESSORS_ONLN); threads = nc > 0 ? (int)nc : 1; }
    if (threads > 1024) threads = 1024;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    pqo_job *jobs = (pqo_job *)malloc(sizeof(pqo_job) * (size_t)threads);
    if (!tid || !jobs) { free(tid); free(jobs); return PQO_ERR_ALLOC; }
    int64_t next = 0; int err = 0, started = 0;
    for (int t = 0; t < threads; ++t) {
        pqo_job j = { c, h, l, v, n_symbols, n_bars, pitch, P, out, ok, &next, 0 };
        jobs[t] = j;
        if (pthread_create(&tid[t], NULL, pqo_worker, &jobs[t]) != 0) break;
        started += 1;
    }
    if (started == 0) { pqo_job j = { c, h, l, v, n_symbols, n_bars, pitch, P, out, ok, &next, 0 }; pqo_worker(&j); err |= j.err; started = 1; }
    else for (int t = 0; t < started; ++t) { pthread_join(tid[t], NULL); err |= jobs[t].err; }
    free(tid); free(jobs);
    return err ? -1 : started;
}
