/* pq_candles.c -- CPU restatement of the reference's candlestick patterns, price transforms and BOP.
 *
 * TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), the cpu_baseline legs of the bench scripts): the
 * product never links or calls this file.
 *
 * Follows src/talib/pattern.rs:9-2065 function by function in the reference's own shape -- one scalar loop per
 * pattern over plain f64 comparisons, `out[i] = +-100`, the helper predicates of pattern.rs:2068-2143 restated
 * one to one -- so that it is independent of the CUDA kernel's formulation (which classifies each bar once into
 * shape flags).  price.rs:10-91 and momentum.rs:113-135 (bop) likewise.  Pinned by tests/golden/candle_golden.npz,
 * the vectors made by executing the reference's pattern.rs text (tests/golden/make_pattern_golden.py): every
 * function below reproduces its golden column exactly (tests/test_candles.py).
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile).  Rust's f64::min / max ignore a NaN operand,
 * like fmin / fmax. */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define EXPORT __attribute__((visibility("default")))

/* ---- pattern.rs:2068-2143 ---- */
static inline int bull(double o, double c) { return c > o; }
static inline int bear(double o, double c) { return c < o; }
static inline double body_abs(double o, double c) { return fabs(o - c); }
static inline double oc_min(double o, double c) { return fmin(o, c); }
static inline double oc_max(double o, double c) { return fmax(o, c); }
static inline double upper_shadow(double o, double h, double c) { return h - oc_max(o, c); }
static inline double lower_shadow(double o, double l, double c) { return oc_min(o, c) - l; }
static inline int long_body(double o, double c) { return body_abs(o, c) > 0.05 * (o + c) * 0.5; }
static inline int short_body(double o, double c) { return body_abs(o, c) < 0.1 * (o + c) * 0.5; }
static inline int doji(double o, double h, double l, double c) { (void)h; (void)l; return body_abs(o, c) <= 0.005 * (o + c) * 0.5; }
static inline int long_up_shadow(double o, double h, double c) { return upper_shadow(o, h, c) > 2.0 * body_abs(o, c); }
static inline int long_dn_shadow(double o, double l, double c) { return lower_shadow(o, l, c) > 2.0 * body_abs(o, c); }
static inline int short_up_shadow(double o, double h, double l, double c) { (void)l; return upper_shadow(o, h, c) < 0.5 * body_abs(o, c); }
static inline int short_dn_shadow(double o, double h, double l, double c) { (void)h; return lower_shadow(o, l, c) < 0.5 * body_abs(o, c); }
static inline int vshort_up_shadow(double o, double h, double l, double c) { (void)l; return upper_shadow(o, h, c) < 0.1 * body_abs(o, c); }
static inline int vshort_dn_shadow(double o, double h, double l, double c) { (void)h; return lower_shadow(o, l, c) < 0.1 * body_abs(o, c); }
static inline int vlong_dn_shadow(double o, double l, double c) { return lower_shadow(o, l, c) > 3.0 * body_abs(o, c); }
static inline int near_(double v1, double v2, double h, double l) { return fabs(v1 - v2) < 0.01 * (h + l) * 0.5; }
static inline int equal_(double v1, double v2, double h, double l) { return fabs(v1 - v2) < 0.001 * (h + l) * 0.5; }

/* every pattern: (open, high, low, close, n, penetration, out) */
#define PAT(name) static void name(const double *open, const double *high, const double *low, const double *close, \
                                   int64_t n, double penetration, int32_t *out)
#define UNUSED (void)open; (void)high; (void)low; (void)close; (void)penetration
/* the bar `k` back from i */
#define O(k) open[i - (k)]
#define H(k) high[i - (k)]
#define L(k) low[i - (k)]
#define C(k) close[i - (k)]
#define SET2(up, dn) do { if (up) out[i] = 100; else if (dn) out[i] = -100; } while (0)

PAT(cdl2crows) { UNUSED;                                                   /* pattern.rs:10 */
    for (int64_t i = 2; i < n; ++i) {
        int bull1 = bull(O(2), C(2)) && long_body(O(2), C(2)), bear2 = bear(O(1), C(1)), gap_up2 = O(1) > C(2);
        int bear3 = bear(O(0), C(0)), open_in2 = (O(0) > O(1)) && (O(0) < C(1)), close_in1 = (C(0) > O(2)) && (C(0) < C(2));
        if (bull1 && bear2 && gap_up2 && bear3 && open_in2 && close_in1) out[i] = -100;
    } }
PAT(cdl3blackcrows) { UNUSED;                                              /* :43 */
    for (int64_t i = 2; i < n; ++i) {
        int b1 = bear(O(2), C(2)) && long_body(O(2), C(2)), b2 = bear(O(1), C(1)) && long_body(O(1), C(1));
        int b3 = bear(O(0), C(0)) && long_body(O(0), C(0));
        int w1 = (O(1) < O(2)) && (O(1) > C(2)), w2 = (O(0) < O(1)) && (O(0) > C(1)), lower = (C(1) < C(2)) && (C(0) < C(1));
        if (b1 && b2 && b3 && w1 && w2 && lower) out[i] = -100;
    } }
PAT(cdl3inside) { UNUSED;                                                  /* :76 */
    for (int64_t i = 2; i < n; ++i) {
        int up = bear(O(2), C(2)) && long_body(O(2), C(2)) && bull(O(1), C(1)) && (C(1) < O(2)) && (O(1) > C(2)) && bull(O(0), C(0)) && (C(0) > O(2));
        int dn = bull(O(2), C(2)) && long_body(O(2), C(2)) && bear(O(1), C(1)) && (O(1) < C(2)) && (C(1) > O(2)) && bear(O(0), C(0)) && (C(0) < O(2));
        SET2(up, dn);
    } }
PAT(cdl3linestrike) { UNUSED;                                              /* :114 */
    for (int64_t i = 3; i < n; ++i) {
        int bull_three = bear(O(3), C(3)) && bear(O(2), C(2)) && bear(O(1), C(1)) && (C(2) < C(3)) && (C(1) < C(2)) &&
                         (O(2) > C(3)) && (O(2) < O(3)) && (O(1) > C(2)) && (O(1) < O(2));
        int bull_strike = bull(O(0), C(0)) && (O(0) < C(1)) && (C(0) > O(3));
        int bear_three = bull(O(3), C(3)) && bull(O(2), C(2)) && bull(O(1), C(1)) && (C(2) > C(3)) && (C(1) > C(2)) &&
                         (O(2) < C(3)) && (O(2) > O(3)) && (O(1) < C(2)) && (O(1) > O(2));
        int bear_strike = bear(O(0), C(0)) && (O(0) > C(1)) && (C(0) < O(3));
        SET2(bull_three && bull_strike, bear_three && bear_strike);
    } }
PAT(cdl3outside) { UNUSED;                                                 /* :160 */
    for (int64_t i = 2; i < n; ++i) {
        int up = bear(O(2), C(2)) && bull(O(1), C(1)) && (O(1) <= C(2)) && (C(1) >= O(2)) && bull(O(0), C(0)) && (C(0) > C(1));
        int dn = bull(O(2), C(2)) && bear(O(1), C(1)) && (O(1) >= C(2)) && (C(1) <= O(2)) && bear(O(0), C(0)) && (C(0) < C(1));
        SET2(up, dn);
    } }
PAT(cdl3starsinsouth) { UNUSED;                                            /* :194 */
    for (int64_t i = 2; i < n; ++i) {
        int bear1 = bear(O(2), C(2)) && long_body(O(2), C(2)), has_ls1 = long_dn_shadow(O(2), L(2), C(2)), bear2 = bear(O(1), C(1));
        int lowerlow2 = L(1) > L(2), higherclose2 = C(1) > C(2), bear3 = bear(O(0), C(0)) && short_body(O(0), C(0));
        int inside3 = (H(0) < H(1)) && (L(0) > L(1));
        if (bear1 && has_ls1 && bear2 && lowerlow2 && higherclose2 && bear3 && inside3) out[i] = 100;
    } }
PAT(cdl3whitesoldiers) { UNUSED;                                           /* :234 */
    for (int64_t i = 2; i < n; ++i) {
        int b1 = bull(O(2), C(2)) && long_body(O(2), C(2)), b2 = bull(O(1), C(1)) && long_body(O(1), C(1)), b3 = bull(O(0), C(0)) && long_body(O(0), C(0));
        int w1 = (O(1) > O(2)) && (O(1) <= C(2)), w2 = (O(0) > O(1)) && (O(0) <= C(1)), higher = (C(1) > C(2)) && (C(0) > C(1));
        if (b1 && b2 && b3 && w1 && w2 && higher) out[i] = 100;
    } }
PAT(cdlabandonedbaby) { UNUSED;                                            /* :268 */
    for (int64_t i = 2; i < n; ++i) {
        int doji2 = doji(O(1), H(1), L(1), C(1));
        int up = bear(O(2), C(2)) && long_body(O(2), C(2)) && doji2 && (H(1) < L(2)) && bull(O(0), C(0)) && (L(0) > H(1));
        int dn = bull(O(2), C(2)) && long_body(O(2), C(2)) && doji2 && (L(1) > H(2)) && bear(O(0), C(0)) && (H(0) < L(1));
        SET2(up, dn);
    } }
PAT(cdladvanceblock) { UNUSED;                                             /* :309 */
    for (int64_t i = 2; i < n; ++i) {
        int bull1 = bull(O(2), C(2)) && long_body(O(2), C(2)), bull2 = bull(O(1), C(1)), bull3 = bull(O(0), C(0));
        int w1 = (O(1) > O(2)) && (O(1) <= C(2)), w2 = (O(0) > O(1)) && (O(0) <= C(1)), higher = (C(1) > C(2)) && (C(0) > C(1));
        int shrinking = body_abs(O(0), C(0)) < body_abs(O(1), C(1));
        if (bull1 && bull2 && bull3 && w1 && w2 && higher && shrinking) out[i] = -100;
    } }
PAT(cdlbelthold) { UNUSED;                                                 /* :345 */
    for (int64_t i = 0; i < n; ++i) {
        int up = bull(O(0), C(0)) && long_body(O(0), C(0)) && vshort_dn_shadow(O(0), H(0), L(0), C(0));
        int dn = bear(O(0), C(0)) && long_body(O(0), C(0)) && vshort_up_shadow(O(0), H(0), L(0), C(0));
        SET2(up, dn);
    } }
PAT(cdlbreakaway) { UNUSED;                                                /* :373 */
    for (int64_t i = 4; i < n; ++i) {
        int up = bear(O(4), C(4)) && long_body(O(4), C(4)) && bear(O(3), C(3)) && (O(3) < C(4)) && (C(2) < C(3)) && bull(O(0), C(0)) &&
                 (C(0) > O(3)) && (C(0) < C(4));
        int dn = bull(O(4), C(4)) && long_body(O(4), C(4)) && bull(O(3), C(3)) && (O(3) > C(4)) && (C(2) > C(3)) && bear(O(0), C(0)) &&
                 (C(0) < O(3)) && (C(0) > C(4));
        SET2(up, dn);
    } }
PAT(cdlclosingmarubozu) { UNUSED;                                          /* :414 */
    for (int64_t i = 0; i < n; ++i) {
        int up = bull(O(0), C(0)) && long_body(O(0), C(0)) && vshort_up_shadow(O(0), H(0), L(0), C(0));
        int dn = bear(O(0), C(0)) && long_body(O(0), C(0)) && vshort_dn_shadow(O(0), H(0), L(0), C(0));
        SET2(up, dn);
    } }
PAT(cdlconcealbabyswall) { UNUSED;                                         /* :442 */
    for (int64_t i = 3; i < n; ++i) {
        int bear1 = bear(O(3), C(3)) && long_body(O(3), C(3));
        int ns1 = vshort_up_shadow(O(3), H(3), L(3), C(3)) && vshort_dn_shadow(O(3), H(3), L(3), C(3));
        int bear2 = bear(O(2), C(2)) && long_body(O(2), C(2));
        int ns2 = vshort_up_shadow(O(2), H(2), L(2), C(2)) && vshort_dn_shadow(O(2), H(2), L(2), C(2));
        int bear3 = bear(O(1), C(1)), high_gap3 = H(1) > C(2), bear4 = bear(O(0), C(0)) && long_body(O(0), C(0));
        int engulf = (O(0) > H(1)) && (C(0) < L(2));
        if (bear1 && ns1 && bear2 && ns2 && (C(2) < C(3)) && bear3 && high_gap3 && bear4 && engulf) out[i] = 100;
    } }
PAT(cdlcounterattack) { UNUSED;                                            /* :487 */
    for (int64_t i = 1; i < n; ++i) {
        int nr = near_(C(0), C(1), H(0), L(0));
        int up = bear(O(1), C(1)) && long_body(O(1), C(1)) && bull(O(0), C(0)) && long_body(O(0), C(0)) && nr;
        int dn = bull(O(1), C(1)) && long_body(O(1), C(1)) && bear(O(0), C(0)) && long_body(O(0), C(0)) && nr;
        SET2(up, dn);
    } }
PAT(cdldarkcloudcover) { UNUSED;                                           /* :519 */
    for (int64_t i = 1; i < n; ++i) {
        int bull1 = bull(O(1), C(1)) && long_body(O(1), C(1)), bear_cur = bear(O(0), C(0)), open_above = O(0) > C(1);
        int close_into = C(0) < (C(1) - (body_abs(O(1), C(1)) * penetration)), close_above = C(0) > O(1);
        if (bull1 && bear_cur && open_above && close_into && close_above) out[i] = -100;
    } }
PAT(cdldoji) { UNUSED;                                                     /* :553 */
    for (int64_t i = 0; i < n; ++i) if (doji(O(0), H(0), L(0), C(0))) out[i] = 100; }
PAT(cdldojistar) { UNUSED;                                                 /* :578 */
    for (int64_t i = 1; i < n; ++i) {
        int doji_cur = doji(O(0), H(0), L(0), C(0));
        double cur_mid = (O(0) + C(0)) / 2.0;
        int up = bear(O(1), C(1)) && long_body(O(1), C(1)) && doji_cur && (cur_mid < C(1));
        int dn = bull(O(1), C(1)) && long_body(O(1), C(1)) && doji_cur && (cur_mid > C(1));
        SET2(up, dn);
    } }
PAT(cdldragonflydoji) { UNUSED;                                            /* :610 */
    for (int64_t i = 0; i < n; ++i)
        if (doji(O(0), H(0), L(0), C(0)) && long_dn_shadow(O(0), L(0), C(0)) && vshort_up_shadow(O(0), H(0), L(0), C(0))) out[i] = 100; }
PAT(cdlengulfing) { UNUSED;                                                /* :635 */
    for (int64_t i = 1; i < n; ++i) {
        int up = bear(O(1), C(1)) && bull(O(0), C(0)) && (O(0) <= C(1)) && (C(0) >= O(1)) && ((O(0) < C(1)) || (C(0) > O(1)));
        int dn = bull(O(1), C(1)) && bear(O(0), C(0)) && (O(0) >= C(1)) && (C(0) <= O(1)) && ((O(0) > C(1)) || (C(0) < O(1)));
        SET2(up, dn);
    } }
PAT(cdleveningdojistar) { UNUSED;                                          /* :665 */
    for (int64_t i = 2; i < n; ++i) {
        int bull1 = bull(O(2), C(2)) && long_body(O(2), C(2)), doji2 = doji(O(1), H(1), L(1), C(1)), gap_up = oc_min(O(1), C(1)) > C(2);
        int bear3 = bear(O(0), C(0)), close_into = C(0) < (C(2) - (body_abs(O(2), C(2)) * penetration));
        if (bull1 && doji2 && gap_up && bear3 && close_into) out[i] = -100;
    } }
PAT(cdleveningstar) { UNUSED;                                              /* :703 */
    for (int64_t i = 2; i < n; ++i) {
        int bull1 = bull(O(2), C(2)) && long_body(O(2), C(2)), short2 = short_body(O(1), C(1)), gap_up = oc_min(O(1), C(1)) > C(2);
        int bear3 = bear(O(0), C(0)), close_into = C(0) < (C(2) - (body_abs(O(2), C(2)) * penetration));
        if (bull1 && short2 && gap_up && bear3 && close_into) out[i] = -100;
    } }
PAT(cdlgapsidesidewhite) { UNUSED;                                         /* :739 */
    for (int64_t i = 2; i < n; ++i) {
        int bull2 = bull(O(1), C(1)), bull3 = bull(O(0), C(0));
        int similar_size = near_(body_abs(O(0), C(0)), body_abs(O(1), C(1)), H(0), L(0)), similaropen = near_(O(0), O(1), H(0), L(0));
        int up_gap = bull(O(2), C(2)) && (O(1) > C(2)) && bull2 && bull3 && similar_size && similaropen;
        int down_gap = bear(O(2), C(2)) && (C(1) < C(2)) && bull2 && bull3 && similar_size && similaropen;
        SET2(up_gap, down_gap);
    } }
PAT(cdlgravestonedoji) { UNUSED;                                           /* :777 */
    for (int64_t i = 0; i < n; ++i)
        if (doji(O(0), H(0), L(0), C(0)) && long_up_shadow(O(0), H(0), C(0)) && vshort_dn_shadow(O(0), H(0), L(0), C(0))) out[i] = -100; }
PAT(cdlhammer) { UNUSED;                                                   /* :802 */
    for (int64_t i = 1; i < n; ++i) {
        double ba = body_abs(O(0), C(0)), ls = lower_shadow(O(0), L(0), C(0));
        int mask = short_body(O(0), C(0)) && (ls > (2.0 * ba)) && vshort_up_shadow(O(0), H(0), L(0), C(0));
        if (mask && bear(O(1), C(1))) out[i] = 100;
    } }
PAT(cdlhangingman) { UNUSED;                                               /* :832 */
    for (int64_t i = 1; i < n; ++i) {
        double ba = body_abs(O(0), C(0)), ls = lower_shadow(O(0), L(0), C(0));
        int mask = short_body(O(0), C(0)) && (ls > (2.0 * ba)) && vshort_up_shadow(O(0), H(0), L(0), C(0));
        if (mask && bull(O(1), C(1))) out[i] = -100;
    } }
PAT(cdlharami) { UNUSED;                                                   /* :862 */
    for (int64_t i = 1; i < n; ++i) {
        int up = bear(O(1), C(1)) && long_body(O(1), C(1)) && bull(O(0), C(0)) && short_body(O(0), C(0)) && (O(0) > C(1)) && (C(0) < O(1));
        int dn = bull(O(1), C(1)) && long_body(O(1), C(1)) && bear(O(0), C(0)) && short_body(O(0), C(0)) && (O(0) < C(1)) && (C(0) > O(1));
        SET2(up, dn);
    } }
PAT(cdlharamicross) { UNUSED;                                              /* :896 */
    for (int64_t i = 1; i < n; ++i) {
        int cur_doji = doji(O(0), H(0), L(0), C(0));
        int up = bear(O(1), C(1)) && long_body(O(1), C(1)) && cur_doji && (oc_max(O(0), C(0)) < O(1)) && (oc_min(O(0), C(0)) > C(1));
        int dn = bull(O(1), C(1)) && long_body(O(1), C(1)) && cur_doji && (oc_max(O(0), C(0)) < C(1)) && (oc_min(O(0), C(0)) > O(1));
        SET2(up, dn);
    } }
PAT(cdlhighwave) { UNUSED;                                                 /* :929 */
    for (int64_t i = 0; i < n; ++i) {
        int mask = short_body(O(0), C(0)) && long_up_shadow(O(0), H(0), C(0)) && long_dn_shadow(O(0), L(0), C(0));
        SET2(mask && bull(O(0), C(0)), mask && bear(O(0), C(0)));
    } }
PAT(cdlhikkake) { UNUSED;                                                  /* :956 */
    for (int64_t i = 2; i < n; ++i) {
        int inside_bar = (H(1) < H(2)) && (L(1) > L(2));
        SET2(inside_bar && (C(0) > H(2)) && bull(O(0), C(0)), inside_bar && (C(0) < L(2)) && bear(O(0), C(0)));
    } }
PAT(cdlhikkakemod) { UNUSED;                                               /* :987 */
    for (int64_t i = 3; i < n; ++i) {
        int inside_bar = (H(2) < H(3)) && (L(2) > L(3)), second_inside = (H(1) < H(2)) && (L(1) > L(2));
        SET2(inside_bar && second_inside && (C(0) > H(3)) && bull(O(0), C(0)), inside_bar && second_inside && (C(0) < L(3)) && bear(O(0), C(0)));
    } }
PAT(cdlhomingpigeon) { UNUSED;                                             /* :1021 */
    for (int64_t i = 1; i < n; ++i)
        if (bear(O(1), C(1)) && long_body(O(1), C(1)) && bear(O(0), C(0)) && short_body(O(0), C(0)) && (O(0) < O(1)) && (C(0) > C(1))) out[i] = 100; }
PAT(cdlidentical3crows) { UNUSED;                                          /* :1048 */
    for (int64_t i = 2; i < n; ++i) {
        int b1 = bear(O(2), C(2)) && long_body(O(2), C(2)), b2 = bear(O(1), C(1)) && long_body(O(1), C(1)), b3 = bear(O(0), C(0)) && long_body(O(0), C(0));
        int eq1 = equal_(O(1), C(2), H(0), L(0)), eq2 = equal_(O(0), C(1), H(0), L(0)), lower = (C(1) < C(2)) && (C(0) < C(1));
        if (b1 && b2 && b3 && eq1 && eq2 && lower) out[i] = -100;
    } }
PAT(cdlinneck) { UNUSED;                                                   /* :1083 */
    for (int64_t i = 1; i < n; ++i)
        if (bear(O(1), C(1)) && long_body(O(1), C(1)) && bull(O(0), C(0)) && (O(0) < C(1)) && near_(C(0), C(1), H(0), L(0))) out[i] = -100; }
PAT(cdlinvertedhammer) { UNUSED;                                           /* :1111 */
    for (int64_t i = 1; i < n; ++i) {
        double ba = body_abs(O(0), C(0)), us = upper_shadow(O(0), H(0), C(0));
        int mask = short_body(O(0), C(0)) && (us > (2.0 * ba)) && vshort_dn_shadow(O(0), H(0), L(0), C(0));
        if (mask && bear(O(1), C(1))) out[i] = 100;
    } }
static inline int marubozu_(double o, double h, double l, double c) { return long_body(o, c) && vshort_up_shadow(o, h, l, c) && vshort_dn_shadow(o, h, l, c); }
PAT(cdlkicking) { UNUSED;                                                  /* :1141 */
    for (int64_t i = 1; i < n; ++i) {
        int m1_bear = bear(O(1), C(1)) && marubozu_(O(1), H(1), L(1), C(1)), m1_bull = bull(O(1), C(1)) && marubozu_(O(1), H(1), L(1), C(1));
        int cur_bull = bull(O(0), C(0)) && marubozu_(O(0), H(0), L(0), C(0)), cur_bear = bear(O(0), C(0)) && marubozu_(O(0), H(0), L(0), C(0));
        SET2(m1_bear && cur_bull && (O(0) > O(1)), m1_bull && cur_bear && (O(0) < O(1)));
    } }
PAT(cdlkickingbylength) { UNUSED;                                          /* :1183 */
    for (int64_t i = 1; i < n; ++i) {
        int m1_bear = bear(O(1), C(1)) && marubozu_(O(1), H(1), L(1), C(1)), m1_bull = bull(O(1), C(1)) && marubozu_(O(1), H(1), L(1), C(1));
        int cur_bull = bull(O(0), C(0)) && marubozu_(O(0), H(0), L(0), C(0)), cur_bear = bear(O(0), C(0)) && marubozu_(O(0), H(0), L(0), C(0));
        double ba1 = body_abs(O(1), C(1)), ba0 = body_abs(O(0), C(0));
        int bull_kick = m1_bear && cur_bull && (O(0) > O(1)), bear_kick = m1_bull && cur_bear && (O(0) < O(1));
        int bull_longer = bull_kick && (ba0 >= ba1), bear_longer = bear_kick && (ba0 >= ba1);
        if (bull_longer || (bull_kick && !bear_longer)) out[i] = 100;
        else if (bear_longer || (bear_kick && !bull_longer)) out[i] = -100;
    } }
PAT(cdlladderbottom) { UNUSED;                                             /* :1229 */
    for (int64_t i = 4; i < n; ++i) {
        int bear1 = bear(O(4), C(4)) && long_body(O(4), C(4)), bear2 = bear(O(3), C(3)) && (C(3) < C(4)), bear3 = bear(O(2), C(2)) && (C(2) < C(3));
        int bear4 = bear(O(1), C(1)), has_upper4 = long_up_shadow(O(1), H(1), C(1)), bull5 = bull(O(0), C(0)) && (O(0) > O(1));
        if (bear1 && bear2 && bear3 && bear4 && has_upper4 && bull5) out[i] = 100;
    } }
PAT(cdllongleggeddoji) { UNUSED;                                           /* :1267 */
    for (int64_t i = 0; i < n; ++i)
        if (doji(O(0), H(0), L(0), C(0)) && long_up_shadow(O(0), H(0), C(0)) && long_dn_shadow(O(0), L(0), C(0))) out[i] = 100; }
PAT(cdllongline) { UNUSED;                                                 /* :1292 */
    for (int64_t i = 0; i < n; ++i) {
        int mask = long_body(O(0), C(0)) && short_up_shadow(O(0), H(0), L(0), C(0)) && short_dn_shadow(O(0), H(0), L(0), C(0));
        SET2(mask && bull(O(0), C(0)), mask && bear(O(0), C(0)));
    } }
PAT(cdlmarubozu) { UNUSED;                                                 /* :1321 */
    for (int64_t i = 0; i < n; ++i) {
        int mask = marubozu_(O(0), H(0), L(0), C(0));
        SET2(mask && bull(O(0), C(0)), mask && bear(O(0), C(0)));
    } }
PAT(cdlmatchinglow) { UNUSED;                                              /* :1349 */
    for (int64_t i = 1; i < n; ++i)
        if (bear(O(1), C(1)) && long_body(O(1), C(1)) && bear(O(0), C(0)) && equal_(C(0), C(1), H(0), L(0))) out[i] = 100; }
PAT(cdlmathold) { UNUSED;                                                  /* :1376 */
    for (int64_t i = 4; i < n; ++i) {
        int bull1 = bull(O(4), C(4)) && long_body(O(4), C(4)), small2 = short_body(O(3), C(3)) && (O(3) > C(4));
        int small3 = short_body(O(2), C(2)), small4 = short_body(O(1), C(1));
        int hold_above = (L(3) > O(4)) && (L(2) > O(4)) && (L(1) > O(4)), bull5 = bull(O(0), C(0)) && (C(0) > C(4));
        if (bull1 && small2 && small3 && small4 && hold_above && bull5) out[i] = 100;
    } }
PAT(cdlmorningdojistar) { UNUSED;                                          /* :1416 */
    for (int64_t i = 2; i < n; ++i) {
        int bear1 = bear(O(2), C(2)) && long_body(O(2), C(2)), doji2 = doji(O(1), H(1), L(1), C(1)), gap_down = oc_max(O(1), C(1)) < C(2);
        int bull3 = bull(O(0), C(0)), close_into = C(0) > (C(2) + (body_abs(O(2), C(2)) * penetration));
        if (bear1 && doji2 && gap_down && bull3 && close_into) out[i] = 100;
    } }
PAT(cdlmorningstar) { UNUSED;                                              /* :1454 */
    for (int64_t i = 2; i < n; ++i) {
        int bear1 = bear(O(2), C(2)) && long_body(O(2), C(2)), short2 = short_body(O(1), C(1)), gap_down = oc_max(O(1), C(1)) < C(2);
        int bull3 = bull(O(0), C(0)), close_into = C(0) > (C(2) + (body_abs(O(2), C(2)) * penetration));
        if (bear1 && short2 && gap_down && bull3 && close_into) out[i] = 100;
    } }
PAT(cdlonneck) { UNUSED;                                                   /* :1490 */
    for (int64_t i = 1; i < n; ++i)
        if (bear(O(1), C(1)) && long_body(O(1), C(1)) && bull(O(0), C(0)) && (O(0) < C(1)) && near_(C(0), L(1), H(0), L(0))) out[i] = -100; }
PAT(cdlpiercing) { UNUSED;                                                 /* :1519 */
    for (int64_t i = 1; i < n; ++i) {
        int bear1 = bear(O(1), C(1)) && long_body(O(1), C(1)), bull_cur = bull(O(0), C(0)), open_below = O(0) < C(1);
        int close_into = C(0) > (C(1) + (body_abs(O(1), C(1)) * penetration)), close_below = C(0) < O(1);
        if (bear1 && bull_cur && open_below && close_into && close_below) out[i] = 100;
    } }
PAT(cdlrickshawman) { UNUSED;                                              /* :1553 */
    for (int64_t i = 0; i < n; ++i) {
        double us = upper_shadow(O(0), H(0), C(0)), ls = lower_shadow(O(0), L(0), C(0));
        if (doji(O(0), H(0), L(0), C(0)) && long_up_shadow(O(0), H(0), C(0)) && long_dn_shadow(O(0), L(0), C(0)) && near_(us, ls, H(0), L(0))) out[i] = 100;
    } }
PAT(cdlrisefall3methods) { UNUSED;                                         /* :1581 */
    for (int64_t i = 4; i < n; ++i) {
        int shorts = short_body(O(3), C(3)) && short_body(O(2), C(2)) && short_body(O(1), C(1));
        int rising = bull(O(4), C(4)) && long_body(O(4), C(4)) && shorts && (H(3) < H(4)) && (H(2) < H(4)) && (H(1) < H(4)) &&
                     (L(3) > L(4)) && (L(2) > L(4)) && (L(1) > L(4)) && bull(O(0), C(0)) && long_body(O(0), C(0)) && (C(0) > C(4));
        int falling = bear(O(4), C(4)) && long_body(O(4), C(4)) && shorts && (L(3) > L(4)) && (L(2) > L(4)) && (L(1) > L(4)) &&
                      (H(3) < H(4)) && (H(2) < H(4)) && (H(1) < H(4)) && bear(O(0), C(0)) && long_body(O(0), C(0)) && (C(0) < C(4));
        SET2(rising, falling);
    } }
PAT(cdlseparatinglines) { UNUSED;                                          /* :1647 */
    for (int64_t i = 1; i < n; ++i) {
        int eq = equal_(O(0), O(1), H(0), L(0));
        int up = bear(O(1), C(1)) && long_body(O(1), C(1)) && bull(O(0), C(0)) && long_body(O(0), C(0)) && eq;
        int dn = bull(O(1), C(1)) && long_body(O(1), C(1)) && bear(O(0), C(0)) && long_body(O(0), C(0)) && eq;
        SET2(up, dn);
    } }
PAT(cdlshootingstar) { UNUSED;                                             /* :1679 */
    for (int64_t i = 1; i < n; ++i) {
        double ba = body_abs(O(0), C(0)), us = upper_shadow(O(0), H(0), C(0));
        int mask = short_body(O(0), C(0)) && (us > (2.0 * ba)) && vshort_dn_shadow(O(0), H(0), L(0), C(0));
        if (mask && bull(O(1), C(1))) out[i] = -100;
    } }
PAT(cdlshortline) { UNUSED;                                                /* :1709 */
    for (int64_t i = 0; i < n; ++i) {
        int mask = short_body(O(0), C(0)) && short_up_shadow(O(0), H(0), L(0), C(0)) && short_dn_shadow(O(0), H(0), L(0), C(0));
        SET2(mask && bull(O(0), C(0)), mask && bear(O(0), C(0)));
    } }
PAT(cdlspinningtop) { UNUSED;                                              /* :1738 */
    for (int64_t i = 0; i < n; ++i) {
        int mask = short_body(O(0), C(0)) && (upper_shadow(O(0), H(0), C(0)) > body_abs(O(0), C(0))) && (lower_shadow(O(0), L(0), C(0)) > body_abs(O(0), C(0)));
        SET2(mask && bull(O(0), C(0)), mask && bear(O(0), C(0)));
    } }
PAT(cdlstalledpattern) { UNUSED;                                           /* :1766 */
    for (int64_t i = 2; i < n; ++i) {
        int bull1 = bull(O(2), C(2)) && long_body(O(2), C(2)), bull2 = bull(O(1), C(1)) && long_body(O(1), C(1)) && (C(1) > C(2));
        int bull3 = bull(O(0), C(0)) && short_body(O(0), C(0)) && (C(0) > C(1)), opens_near = (O(0) > O(1)) && (O(0) <= C(1));
        if (bull1 && bull2 && bull3 && opens_near) out[i] = -100;
    } }
PAT(cdlsticksandwich) { UNUSED;                                            /* :1797 */
    for (int64_t i = 2; i < n; ++i)
        if (bear(O(2), C(2)) && long_body(O(2), C(2)) && bull(O(1), C(1)) && long_body(O(1), C(1)) && (O(1) > C(2)) && bear(O(0), C(0)) &&
            long_body(O(0), C(0)) && equal_(C(0), C(2), H(0), L(0))) out[i] = 100; }
PAT(cdltakuri) { UNUSED;                                                   /* :1831 */
    for (int64_t i = 0; i < n; ++i)
        if (doji(O(0), H(0), L(0), C(0)) && vlong_dn_shadow(O(0), L(0), C(0)) && vshort_up_shadow(O(0), H(0), L(0), C(0))) out[i] = 100; }
PAT(cdltasukigap) { UNUSED;                                                /* :1856 */
    for (int64_t i = 2; i < n; ++i) {
        int up = bull(O(2), C(2)) && bull(O(1), C(1)) && (O(1) > C(2)) && bear(O(0), C(0)) && (O(0) > O(1)) && (O(0) < C(1)) && (C(0) > O(2)) && (C(0) < C(2));
        int dn = bear(O(2), C(2)) && bear(O(1), C(1)) && (O(1) < C(2)) && bull(O(0), C(0)) && (O(0) < O(1)) && (O(0) > C(1)) && (C(0) < O(2)) && (C(0) > C(2));
        SET2(up, dn);
    } }
PAT(cdlthrusting) { UNUSED;                                                /* :1894 */
    for (int64_t i = 1; i < n; ++i) {
        double midpoint = C(1) + (body_abs(O(1), C(1)) * 0.5);
        if (bear(O(1), C(1)) && long_body(O(1), C(1)) && bull(O(0), C(0)) && (O(0) < C(1)) && (C(0) > C(1)) && (C(0) < midpoint)) out[i] = -100;
    } }
PAT(cdltristar) { UNUSED;                                                  /* :1922 */
    for (int64_t i = 2; i < n; ++i) {
        int d3 = doji(O(2), H(2), L(2), C(2)) && doji(O(1), H(1), L(1), C(1)) && doji(O(0), H(0), L(0), C(0));
        double mid1 = (O(2) + C(2)) / 2.0, mid2 = (O(1) + C(1)) / 2.0, mid3 = (O(0) + C(0)) / 2.0;
        SET2(d3 && (mid2 < mid1) && (mid3 > mid2), d3 && (mid2 > mid1) && (mid3 < mid2));
    } }
PAT(cdlunique3river) { UNUSED;                                             /* :1964 */
    for (int64_t i = 2; i < n; ++i) {
        int bear1 = bear(O(2), C(2)) && long_body(O(2), C(2)), bear2 = bear(O(1), C(1)) && (L(1) < L(2)) && (C(1) > L(1));
        int harami = (O(1) < O(2)) && (O(1) > C(2)), bull3 = bull(O(0), C(0)) && short_body(O(0), C(0)) && (C(0) < C(1));
        if (bear1 && bear2 && harami && bull3) out[i] = 100;
    } }
PAT(cdlupsidegap2crows) { UNUSED;                                          /* :1997 */
    for (int64_t i = 2; i < n; ++i) {
        int bull1 = bull(O(2), C(2)) && long_body(O(2), C(2)), bear2 = bear(O(1), C(1)) && (O(1) > C(2)) && (C(1) > C(2));
        int bear3 = bear(O(0), C(0)) && (O(0) > O(1)) && (C(0) > C(2)) && (C(0) < C(1));
        if (bull1 && bear2 && bear3) out[i] = -100;
    } }
PAT(cdlxsidegap3methods) { UNUSED;                                         /* :2027 */
    for (int64_t i = 2; i < n; ++i) {
        int up = bull(O(2), C(2)) && bull(O(1), C(1)) && (O(1) > C(2)) && bear(O(0), C(0)) && (O(0) < C(1)) && (O(0) > O(1)) && (C(0) > O(2)) && (C(0) < C(2));
        int dn = bear(O(2), C(2)) && bear(O(1), C(1)) && (O(1) < C(2)) && bull(O(0), C(0)) && (O(0) > C(1)) && (O(0) < O(1)) && (C(0) < O(2)) && (C(0) > C(2));
        SET2(up, dn);
    } }

typedef void (*pat_fn)(const double *, const double *, const double *, const double *, int64_t, double, int32_t *);
static const pat_fn PATTERNS[61] = {                /* the reference's order of definition */
    cdl2crows, cdl3blackcrows, cdl3inside, cdl3linestrike, cdl3outside, cdl3starsinsouth, cdl3whitesoldiers, cdlabandonedbaby,
    cdladvanceblock, cdlbelthold, cdlbreakaway, cdlclosingmarubozu, cdlconcealbabyswall, cdlcounterattack, cdldarkcloudcover,
    cdldoji, cdldojistar, cdldragonflydoji, cdlengulfing, cdleveningdojistar, cdleveningstar, cdlgapsidesidewhite,
    cdlgravestonedoji, cdlhammer, cdlhangingman, cdlharami, cdlharamicross, cdlhighwave, cdlhikkake, cdlhikkakemod,
    cdlhomingpigeon, cdlidentical3crows, cdlinneck, cdlinvertedhammer, cdlkicking, cdlkickingbylength, cdlladderbottom,
    cdllongleggeddoji, cdllongline, cdlmarubozu, cdlmatchinglow, cdlmathold, cdlmorningdojistar, cdlmorningstar, cdlonneck,
    cdlpiercing, cdlrickshawman, cdlrisefall3methods, cdlseparatinglines, cdlshootingstar, cdlshortline, cdlspinningtop,
    cdlstalledpattern, cdlsticksandwich, cdltakuri, cdltasukigap, cdlthrusting, cdltristar, cdlunique3river, cdlupsidegap2crows,
    cdlxsidegap3methods};

/* one pattern on one symbol's columns (out is zeroed first, like `vec![0i32; n]`) */
EXPORT int pqc_pattern(int pattern, const double *open, const double *high, const double *low, const double *close, int64_t n,
                       double penetration, int32_t *out) {
    if (pattern < 0 || pattern >= 61 || n < 0) return -1;
    memset(out, 0, (size_t)n * sizeof(int32_t));
    PATTERNS[pattern](open, high, low, close, n, penetration, out);
    return 0;
}

/* price.rs:10-91, momentum.rs:113-135 (null-free columns) */
EXPORT int pqc_price(int which, const double *o, const double *h, const double *l, const double *c, int64_t n, double *out) {
    for (int64_t i = 0; i < n; ++i) {
        switch (which) {
            case 0: out[i] = (o[i] + h[i] + l[i] + c[i]) * 0.25; break;
            case 1: out[i] = (h[i] + l[i]) * 0.5; break;
            case 2: out[i] = (h[i] + l[i] + c[i]) / 3.0; break;
            case 3: out[i] = (h[i] + l[i] + 2.0 * c[i]) / 4.0; break;
            case 4: { double diff = h[i] - l[i]; out[i] = (diff == 0.0) ? 0.0 : (c[i] - o[i]) / diff; break; }
            default: return -1;
        }
    }
    return 0;
}

/* the whole candle family over a row-major panel, `threads` workers over symbols: how the reference would run it
 * (66 plugin calls per symbol, each its own pass over the columns) -- the CPU timing leg */
typedef struct {
    const double *o, *h, *l, *c; int64_t n_symbols, n_bars, pitch; double pen; int32_t *pat; double *price; int64_t *next;
} cjob;
static void *cworker(void *arg) {
    cjob *J = (cjob *)arg;
    const size_t plane = (size_t)J->n_symbols * (size_t)J->pitch;
    for (;;) {
        int64_t s = __sync_fetch_and_add(J->next, 1);
        if (s >= J->n_symbols) break;
        const size_t r = (size_t)s * (size_t)J->pitch;
        for (int k = 0; k < 61; ++k) pqc_pattern(k, J->o + r, J->h + r, J->l + r, J->c + r, J->n_bars, J->pen, J->pat + k * plane + r);
        for (int k = 0; k < 5; ++k) pqc_price(k, J->o + r, J->h + r, J->l + r, J->c + r, J->n_bars, J->price + k * plane + r);
    }
    return NULL;
}
EXPORT int pqc_panel(const double *o, const double *h, const double *l, const double *c, int64_t n_symbols, int64_t n_bars,
                     int64_t pitch, double penetration, int32_t *pat /* [61][S][pitch] */, double *price /* [5][S][pitch] */,
                     int threads) {
    if (threads <= 0) { long nc = sysconf(_SC_NPROCESSORS_ONLN); threads = nc > 0 ? (int)nc : 1; }
    if (threads > 1024) threads = 1024;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    if (!tid) return -1;
    int64_t next = 0;
    cjob J = {o, h, l, c, n_symbols, n_bars, pitch, penetration, pat, price, &next};
    int started = 0;
    for (int t = 0; t < threads; ++t) { if (pthread_create(&tid[t], NULL, cworker, &J) != 0) break; started += 1; }
    if (started == 0) { cworker(&J); started = 1; }
    else for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
    free(tid);
    return started;
}
