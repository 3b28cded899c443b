// candles.cuh -- fused candle kernel for sm_100a: the reference's 61 candlestick patterns
// (src/talib/pattern.rs:9-2065, Int32 outputs in {-100, 0, 100}), the four price transforms
// (src/talib/price.rs:10-91) and BOP (src/talib/momentum.rs:113-135) from open / high / low / close
// in ONE pass.
//
// The reference runs 66 separate plugin calls per symbol, each re-reading the four columns and each
// re-deriving the same per-bar quantities (body, shadows, long / short / doji tests) for up to five
// bars.  Here a bar's shape is classified ONCE into a 16-bit flag word; every pattern is then a
// boolean combination of the flag words of bars t .. t-4 plus a few cross-bar comparisons of raw
// prices, so the whole family costs ~35 FP64 operations per bar instead of ~600.
//
// Layout: row-major [symbol][pitch] planes -- the Arrow-shaped layout of the ABI -- because this is a
// pure stencil: a warp takes 32 consecutive bars of one symbol (256 B per input plane, 128 B per Int32
// output plane, whole cache lines either way), a CTA 256 consecutive bars (+4 halo bars) staged in
// shared memory.  HBM-bound: 32 B in + 4 B x patterns + 8 B x prices per symbol-bar (316 B with
// everything enabled), every output written exactly once with streaming stores.
// Arithmetic is f64 in the reference's expression order (-fmad=false), comparisons only: results
// are bit-exact.  Nulls: patterns and BOP refuse them on the host like the reference's cont_slice()?;
// the price transforms propagate them (validity word = AND of the inputs' words, one ballot per warp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pqb {

// pattern ids: the reference's (alphabetical) order of definition in pattern.rs
#define PQB_PATTERN_LIST(X)                                                                                  \
    X(cdl2crows) X(cdl3blackcrows) X(cdl3inside) X(cdl3linestrike) X(cdl3outside) X(cdl3starsinsouth)        \
    X(cdl3whitesoldiers) X(cdlabandonedbaby) X(cdladvanceblock) X(cdlbelthold) X(cdlbreakaway)               \
    X(cdlclosingmarubozu) X(cdlconcealbabyswall) X(cdlcounterattack) X(cdldarkcloudcover) X(cdldoji)         \
    X(cdldojistar) X(cdldragonflydoji) X(cdlengulfing) X(cdleveningdojistar) X(cdleveningstar)               \
    X(cdlgapsidesidewhite) X(cdlgravestonedoji) X(cdlhammer) X(cdlhangingman) X(cdlharami) X(cdlharamicross) \
    X(cdlhighwave) X(cdlhikkake) X(cdlhikkakemod) X(cdlhomingpigeon) X(cdlidentical3crows) X(cdlinneck)      \
    X(cdlinvertedhammer) X(cdlkicking) X(cdlkickingbylength) X(cdlladderbottom) X(cdllongleggeddoji)         \
    X(cdllongline) X(cdlmarubozu) X(cdlmatchinglow) X(cdlmathold) X(cdlmorningdojistar) X(cdlmorningstar)    \
    X(cdlonneck) X(cdlpiercing) X(cdlrickshawman) X(cdlrisefall3methods) X(cdlseparatinglines)               \
    X(cdlshootingstar) X(cdlshortline) X(cdlspinningtop) X(cdlstalledpattern) X(cdlsticksandwich)            \
    X(cdltakuri) X(cdltasukigap) X(cdlthrusting) X(cdltristar) X(cdlunique3river) X(cdlupsidegap2crows)      \
    X(cdlxsidegap3methods)

enum PatternId {
#define X(n) P_##n,
    PQB_PATTERN_LIST(X)
#undef X
    N_PATTERNS
};
static_assert(N_PATTERNS == 61, "pattern.rs defines 61 functions");

constexpr int N_PRICES = 5;          // avgprice, medprice, typprice, wclprice, bop
constexpr int CANDLE_TILE = 256;     // bars per CTA
constexpr int CANDLE_HALO = 4;       // deepest lookback (breakaway, ladderbottom, mathold, risefall3methods)

// the six functions that read a `penetration` literal (pattern.rs `inputs.get(4) ... unwrap_or(0.3)`)
enum { PEN_DARKCLOUDCOVER, PEN_EVENINGDOJISTAR, PEN_EVENINGSTAR, PEN_MORNINGDOJISTAR, PEN_MORNINGSTAR, PEN_PIERCING, N_PEN };

struct CandleArgs {
    const double *in[4];             // open, high, low, close: row-major [n_symbols][pitch]
    const uint32_t *vin[4];          // input validity words [n_symbols][words_per_row] or nullptr (all valid)
    int32_t *pat[N_PATTERNS];        // Int32 planes [n_symbols][pitch] or nullptr
    double *price[N_PRICES];         // f64 planes or nullptr
    uint32_t *vprice[N_PRICES];      // validity words of the price planes or nullptr
    double pen[N_PEN];
    size_t pat_stride;               // ALL specialisation: pattern k's plane = pat[0] + k * pat_stride (one allocation)
    int n_symbols, n_bars, pitch, words_per_row;
    int symbol0;                     // first symbol of this launch (chunked host pipeline)
};

// per-bar shape flags (pattern.rs:2068-2143 helper predicates, each evaluated once per bar)
enum : unsigned {
    CF_BULL = 1u << 0,    // c > o
    CF_BEAR = 1u << 1,    // c < o
    CF_LONG = 1u << 2,    // |o-c| > 0.05 * (o+c) * 0.5
    CF_SHORT = 1u << 3,   // |o-c| < 0.1 * (o+c) * 0.5
    CF_DOJI = 1u << 4,    // |o-c| <= 0.005 * (o+c) * 0.5
    CF_LUS = 1u << 5,     // upper shadow > 2 body
    CF_LDS = 1u << 6,     // lower shadow > 2 body
    CF_SUS = 1u << 7,     // upper shadow < 0.5 body
    CF_SDS = 1u << 8,     // lower shadow < 0.5 body
    CF_VSUS = 1u << 9,    // upper shadow < 0.1 body
    CF_VSDS = 1u << 10,   // lower shadow < 0.1 body
    CF_VLDS = 1u << 11,   // lower shadow > 3 body
    CF_USGB = 1u << 12,   // upper shadow > body   (spinningtop)
    CF_LSGB = 1u << 13,   // lower shadow > body
};

__device__ __forceinline__ unsigned classify(double o, double h, double l, double c) {
    const double body = fabs(o - c);
    const double sum = o + c;
    const double us = h - fmax(o, c);          // Rust f64::max / min: a NaN operand is ignored, like fmax / fmin
    const double ls = fmin(o, c) - l;
    unsigned f = 0;
    f |= (c > o) ? CF_BULL : 0u;
    f |= (c < o) ? CF_BEAR : 0u;
    f |= (body > 0.05 * sum * 0.5) ? CF_LONG : 0u;
    f |= (body < 0.1 * sum * 0.5) ? CF_SHORT : 0u;
    f |= (body <= 0.005 * sum * 0.5) ? CF_DOJI : 0u;
    const double b2 = 2.0 * body, bh = 0.5 * body, bt = 0.1 * body;
    f |= (us > b2) ? CF_LUS : 0u;
    f |= (ls > b2) ? CF_LDS : 0u;
    f |= (us < bh) ? CF_SUS : 0u;
    f |= (ls < bh) ? CF_SDS : 0u;
    f |= (us < bt) ? CF_VSUS : 0u;
    f |= (ls < bt) ? CF_VSDS : 0u;
    f |= (ls > 3.0 * body) ? CF_VLDS : 0u;
    f |= (us > body) ? CF_USGB : 0u;
    f |= (ls > body) ? CF_LSGB : 0u;
    return f;
}

__device__ __forceinline__ void st_i32(int32_t *p, int v) { __stcs(p, v); }

// grid = (ceil(n_bars / CANDLE_TILE), symbols of this launch), CANDLE_TILE threads.
// ALL: every pattern enabled and the 61 planes are one allocation (no mask tests, one base pointer).
template <bool ALL>
__global__ void __launch_bounds__(CANDLE_TILE, 3) candle_kernel(const __grid_constant__ CandleArgs A, const uint64_t pmask) {
    __shared__ double so[CANDLE_TILE + CANDLE_HALO], sh[CANDLE_TILE + CANDLE_HALO], sl[CANDLE_TILE + CANDLE_HALO],
        sc[CANDLE_TILE + CANDLE_HALO];
    __shared__ unsigned short sf[CANDLE_TILE + CANDLE_HALO];
    const int tid = threadIdx.x;
    const int s = A.symbol0 + blockIdx.y;
    const int t0 = blockIdx.x * CANDLE_TILE;
    const int i = t0 + tid;                                  // this thread's bar
    const size_t row = (size_t)s * A.pitch;
    const bool live = i < A.n_bars;

    // ---- stage the tile (+ halo) and classify every bar once
    double o0 = 0.0, h0 = 0.0, l0 = 0.0, c0 = 0.0;
    if (live) {
        o0 = __ldg(A.in[0] + row + i);
        h0 = __ldg(A.in[1] + row + i);
        l0 = __ldg(A.in[2] + row + i);
        c0 = __ldg(A.in[3] + row + i);
    }
    const unsigned f0 = classify(o0, h0, l0, c0);
    so[tid + CANDLE_HALO] = o0; sh[tid + CANDLE_HALO] = h0; sl[tid + CANDLE_HALO] = l0; sc[tid + CANDLE_HALO] = c0;
    sf[tid + CANDLE_HALO] = (unsigned short)f0;
    if (tid < CANDLE_HALO) {
        const int j = t0 - CANDLE_HALO + tid;
        double o = 0.0, h = 0.0, l = 0.0, c = 0.0;
        if (j >= 0) {
            o = __ldg(A.in[0] + row + j);
            h = __ldg(A.in[1] + row + j);
            l = __ldg(A.in[2] + row + j);
            c = __ldg(A.in[3] + row + j);
        }
        so[tid] = o; sh[tid] = h; sl[tid] = l; sc[tid] = c;
        sf[tid] = (unsigned short)classify(o, h, l, c);
    }
    __syncthreads();

    // ---- price transforms + BOP (null-propagating: validity = AND of the inputs' words; one word per warp)
    {
        const int w = i >> 5;                                // validity word of this warp's 32 bars (t0 is a multiple of 32)
        unsigned vo = 0xffffffffu, vh = vo, vl = vo, vc = vo;
        if (live) {
            const size_t wi = (size_t)s * A.words_per_row + w;
            if (A.vin[0]) vo = A.vin[0][wi];
            if (A.vin[1]) vh = A.vin[1][wi];
            if (A.vin[2]) vl = A.vin[2][wi];
            if (A.vin[3]) vc = A.vin[3][wi];
        }
        const unsigned bit = 1u << (tid & 31);
        const bool ko = vo & bit, kh = vh & bit, kl = vl & bit, kc = vc & bit;
        const double nn = __longlong_as_double(0x7ff8000000000000LL);
        auto emit = [&](int k, double v, bool ok) {
            if (!A.price[k]) return;
            ok = ok && live;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (live) {
                __stcs(A.price[k] + row + i, ok ? v : nn);
                if ((tid & 31) == 0 && A.vprice[k]) A.vprice[k][(size_t)s * A.words_per_row + w] = m;
            }
        };
        emit(0, (o0 + h0 + l0 + c0) * 0.25, ko && kh && kl && kc);          // avgprice price.rs:10
        emit(1, (h0 + l0) * 0.5, kh && kl);                                 // medprice :35
        emit(2, (h0 + l0 + c0) / 3.0, kh && kl && kc);                      // typprice :55
        emit(3, (h0 + l0 + 2.0 * c0) / 4.0, kh && kl && kc);                // wclprice :75
        const double diff = h0 - l0;                                        // bop momentum.rs:113-135
        emit(4, (diff == 0.0) ? 0.0 : (c0 - o0) / ((diff == 0.0) ? 1.0 : diff), ko && kh && kl && kc);
    }
    if (!ALL && pmask == 0) return;

    // ---- the window: k bars back = index tid + HALO - k
    const int q = tid + CANDLE_HALO;
    const double o1 = so[q - 1], h1 = sh[q - 1], l1 = sl[q - 1], c1 = sc[q - 1];
    const double o2 = so[q - 2], h2 = sh[q - 2], l2 = sl[q - 2], c2 = sc[q - 2];
    const double o3 = so[q - 3], h3 = sh[q - 3], l3 = sl[q - 3], c3 = sc[q - 3];
    const double o4 = so[q - 4], h4 = sh[q - 4], l4 = sl[q - 4], c4 = sc[q - 4];
    const unsigned f1 = sf[q - 1], f2 = sf[q - 2], f3 = sf[q - 3], f4 = sf[q - 4];
    const double b0 = fabs(o0 - c0), b1 = fabs(o1 - c1), b2 = fabs(o2 - c2);
    const double us0 = h0 - fmax(o0, c0), ls0 = fmin(o0, c0) - l0;
    const double hl0 = h0 + l0;
    const double near_t = 0.01 * hl0 * 0.5, equal_t = 0.001 * hl0 * 0.5;     // near / equal use the CURRENT bar's range
#define HAS(f, m) (((f) & (m)) == (m))
#define NEAR(a, b) (fabs((a) - (b)) < near_t)
#define EQUAL(a, b) (fabs((a) - (b)) < equal_t)
    const bool bull0 = f0 & CF_BULL, bear0 = f0 & CF_BEAR, bull1 = f1 & CF_BULL, bear1 = f1 & CF_BEAR;
    const bool bull2 = f2 & CF_BULL, bear2 = f2 & CF_BEAR;
    const bool lbull0 = HAS(f0, CF_BULL | CF_LONG), lbear0 = HAS(f0, CF_BEAR | CF_LONG);
    const bool lbull1 = HAS(f1, CF_BULL | CF_LONG), lbear1 = HAS(f1, CF_BEAR | CF_LONG);
    const bool lbull2 = HAS(f2, CF_BULL | CF_LONG), lbear2 = HAS(f2, CF_BEAR | CF_LONG);
    const bool lbull4 = HAS(f4, CF_BULL | CF_LONG), lbear4 = HAS(f4, CF_BEAR | CF_LONG);
    const bool doji0 = f0 & CF_DOJI, doji1 = f1 & CF_DOJI, short0 = f0 & CF_SHORT, short1 = f1 & CF_SHORT;
    const bool maru0 = HAS(f0, CF_LONG | CF_VSUS | CF_VSDS), maru1 = HAS(f1, CF_LONG | CF_VSUS | CF_VSDS);
    const bool lb1 = i >= 1, lb2 = i >= 2, lb3 = i >= 3, lb4 = i >= 4;
    const size_t at = row + i;
    int32_t *const pbase = A.pat[0] + at;
    const size_t pstride = A.pat_stride;
    // `up` -> +100, else `dn` -> -100, else 0.  (A warp-voted early-out on the flag preconditions was measured:
    // on random-walk data almost every 32-bar warp has a lane that passes, and the votes cost more than they save.)
#define EMIT(id, up, dn)                                                                   \
    if (ALL || (pmask >> (id) & 1)) {                                                      \
        if (live) st_i32(ALL ? pbase + (size_t)(id) * pstride : A.pat[id] + at, (up) ? 100 : ((dn) ? -100 : 0)); \
    }
    const bool bear3 = f3 & CF_BEAR, bull3 = f3 & CF_BULL, short2 = f2 & CF_SHORT, short3 = f3 & CF_SHORT, doji2 = f2 & CF_DOJI;
    const bool long0 = f0 & CF_LONG, long1 = f1 & CF_LONG, long2 = f2 & CF_LONG, long4 = f4 & CF_LONG;

    EMIT(P_cdl2crows, false,                                      // pattern.rs:10 (can never fire:
         lb2 && lbull2 && bear1 && o1 > c2 && bear0 && (o0 > o1 && o0 < c1) && (c0 > o2 && c0 < c2))   //  o0 > o1 && o0 < c1 with c1 < o1)
    EMIT(P_cdl3blackcrows, false,                               // :43
         lb2 && lbear2 && lbear1 && lbear0 && (o1 < o2 && o1 > c2) && (o0 < o1 && o0 > c1) && (c1 < c2 && c0 < c1))
    EMIT(P_cdl3inside,                                                               // :76
         lb2 && lbear2 && bull1 && c1 < o2 && o1 > c2 && bull0 && c0 > o2,
         lb2 && lbull2 && bear1 && o1 < c2 && c1 > o2 && bear0 && c0 < o2)
    EMIT(P_cdl3linestrike,   // :114
         lb3 && bear3 && bear2 && bear1 && c2 < c3 && c1 < c2 && o2 > c3 && o2 < o3 && o1 > c2 && o1 < o2 && bull0 && o0 < c1 && c0 > o3,
         lb3 && bull3 && bull2 && bull1 && c2 > c3 && c1 > c2 && o2 < c3 && o2 > o3 && o1 < c2 && o1 > o2 && bear0 && o0 > c1 && c0 < o3)
    EMIT(P_cdl3outside,           // :160
         lb2 && bear2 && bull1 && o1 <= c2 && c1 >= o2 && bull0 && c0 > c1,
         lb2 && bull2 && bear1 && o1 >= c2 && c1 <= o2 && bear0 && c0 < c1)
    EMIT(P_cdl3starsinsouth,           // :194
         lb2 && lbear2 && (f2 & CF_LDS) && bear1 && l1 > l2 && c1 > c2 && bear0 && short0 && h0 < h1 && l0 > l1, false)
    EMIT(P_cdl3whitesoldiers,                                   // :234
         lb2 && lbull2 && lbull1 && lbull0 && (o1 > o2 && o1 <= c2) && (o0 > o1 && o0 <= c1) && (c1 > c2 && c0 > c1), false)
    EMIT(P_cdlabandonedbaby,                                                // :268
         lb2 && lbear2 && doji1 && h1 < l2 && bull0 && l0 > h1,
         lb2 && lbull2 && doji1 && l1 > h2 && bear0 && h0 < l1)
    EMIT(P_cdladvanceblock, false,                                // :309
         lb2 && lbull2 && bull1 && bull0 && (o1 > o2 && o1 <= c2) && (o0 > o1 && o0 <= c1) && (c1 > c2 && c0 > c1) && b0 < b1)
    EMIT(P_cdlbelthold, lbull0 && (f0 & CF_VSDS), lbear0 && (f0 & CF_VSUS))                       // :345
    EMIT(P_cdlbreakaway,                                                             // :373
         lb4 && lbear4 && bear3 && o3 < c4 && c2 < c3 && bull0 && c0 > o3 && c0 < c4,
         lb4 && lbull4 && bull3 && o3 > c4 && c2 > c3 && bear0 && c0 < o3 && c0 > c4)
    EMIT(P_cdlclosingmarubozu, lbull0 && (f0 & CF_VSUS), lbear0 && (f0 & CF_VSDS))                // :414
    EMIT(P_cdlconcealbabyswall,
         lb3 && HAS(f3, CF_BEAR | CF_LONG | CF_VSUS | CF_VSDS) && HAS(f2, CF_BEAR | CF_LONG | CF_VSUS | CF_VSDS) && c2 < c3 &&
             bear1 && h1 > c2 && lbear0 && o0 > h1 && c0 < l2, false)
    EMIT(P_cdlcounterattack,                                                // :487
         lb1 && lbear1 && lbull0 && NEAR(c0, c1), lb1 && lbull1 && lbear0 && NEAR(c0, c1))
    EMIT(P_cdldarkcloudcover, false,                                       // :519
         lb1 && lbull1 && bear0 && o0 > c1 && c0 < (c1 - (b1 * A.pen[PEN_DARKCLOUDCOVER])) && c0 > o1)
    EMIT(P_cdldoji, doji0, false)                                                                 // :553
    EMIT(P_cdldojistar,                                                     // :578
         lb1 && lbear1 && doji0 && ((o0 + c0) / 2.0) < c1, lb1 && lbull1 && doji0 && ((o0 + c0) / 2.0) > c1)
    EMIT(P_cdldragonflydoji, HAS(f0, CF_DOJI | CF_LDS | CF_VSUS), false)                          // :610
    EMIT(P_cdlengulfing,                            // :635
         lb1 && bear1 && bull0 && o0 <= c1 && c0 >= o1 && (o0 < c1 || c0 > o1),
         lb1 && bull1 && bear0 && o0 >= c1 && c0 <= o1 && (o0 > c1 || c0 < o1))
    EMIT(P_cdleveningdojistar, false,                             // :665
         lb2 && lbull2 && doji1 && fmin(o1, c1) > c2 && bear0 && c0 < (c2 - (b2 * A.pen[PEN_EVENINGDOJISTAR])))
    EMIT(P_cdleveningstar, false,                                // :703
         lb2 && lbull2 && short1 && fmin(o1, c1) > c2 && bear0 && c0 < (c2 - (b2 * A.pen[PEN_EVENINGSTAR])))
    EMIT(P_cdlgapsidesidewhite,                                             // :739
         lb2 && bull2 && o1 > c2 && bull1 && bull0 && NEAR(b0, b1) && NEAR(o0, o1),
         lb2 && bear2 && c1 < c2 && bull1 && bull0 && NEAR(b0, b1) && NEAR(o0, o1))
    EMIT(P_cdlgravestonedoji, false, HAS(f0, CF_DOJI | CF_LUS | CF_VSDS))                         // :777
    EMIT(P_cdlhammer, lb1 && HAS(f0, CF_SHORT | CF_LDS | CF_VSUS) && bear1, false)                // :802
    EMIT(P_cdlhangingman, false, lb1 && HAS(f0, CF_SHORT | CF_LDS | CF_VSUS) && bull1)            // :832
    EMIT(P_cdlharami,                                                      // :862
         lb1 && lbear1 && bull0 && short0 && o0 > c1 && c0 < o1,
         lb1 && lbull1 && bear0 && short0 && o0 < c1 && c0 > o1)
    EMIT(P_cdlharamicross,                                                  // :896
         lb1 && lbear1 && doji0 && fmax(o0, c0) < o1 && fmin(o0, c0) > c1,
         lb1 && lbull1 && doji0 && fmax(o0, c0) < c1 && fmin(o0, c0) > o1)
    EMIT(P_cdlhighwave, HAS(f0, CF_SHORT | CF_LUS | CF_LDS) && bull0, HAS(f0, CF_SHORT | CF_LUS | CF_LDS) && bear0)   // :929
    {                                                                                              // :956, :987 hikkake / hikkakemod
        const bool inside12 = h1 < h2 && l1 > l2;
        EMIT(P_cdlhikkake, lb2 && inside12 && c0 > h2 && bull0, lb2 && inside12 && c0 < l2 && bear0)
        EMIT(P_cdlhikkakemod,
             lb3 && (h2 < h3 && l2 > l3) && inside12 && c0 > h3 && bull0, lb3 && (h2 < h3 && l2 > l3) && inside12 && c0 < l3 && bear0)
    }
    EMIT(P_cdlhomingpigeon,                                      // :1021
         lb1 && lbear1 && bear0 && short0 && o0 < o1 && c0 > c1, false)
    EMIT(P_cdlidentical3crows, false,                           // :1048
         lb2 && lbear2 && lbear1 && lbear0 && EQUAL(o1, c2) && EQUAL(o0, c1) && (c1 < c2 && c0 < c1))
    EMIT(P_cdlinneck, false, lb1 && lbear1 && bull0 && o0 < c1 && NEAR(c0, c1))   // :1083
    EMIT(P_cdlinvertedhammer, lb1 && HAS(f0, CF_SHORT | CF_LUS | CF_VSDS) && bear1, false)        // :1111
    {                                                                                              // :1141, :1183 kicking / kickingbylength
        const bool pre = lb1 && maru1 && maru0;
        const bool bull_kick = pre && bear1 && bull0 && o0 > o1;
        const bool bear_kick = pre && bull1 && bear0 && o0 < o1;
        EMIT(P_cdlkicking, bull_kick, bear_kick)
        EMIT(P_cdlkickingbylength, (bull_kick && b0 >= b1) || (bull_kick && !(bear_kick && b0 >= b1)),
             (bear_kick && b0 >= b1) || (bear_kick && !(bull_kick && b0 >= b1)))
    }
    EMIT(P_cdlladderbottom,    // :1229
         lb4 && lbear4 && (bear3 && c3 < c4) && (bear2 && c2 < c3) && bear1 && (f1 & CF_LUS) && (bull0 && o0 > o1), false)
    EMIT(P_cdllongleggeddoji, HAS(f0, CF_DOJI | CF_LUS | CF_LDS), false)                          // :1267
    EMIT(P_cdllongline, HAS(f0, CF_LONG | CF_SUS | CF_SDS) && bull0, HAS(f0, CF_LONG | CF_SUS | CF_SDS) && bear0)    // :1292
    EMIT(P_cdlmarubozu, maru0 && bull0, maru0 && bear0)                                           // :1321
    EMIT(P_cdlmatchinglow, lb1 && lbear1 && bear0 && EQUAL(c0, c1), false) // :1349
    EMIT(P_cdlmathold,                       // :1376
         lb4 && lbull4 && (short3 && o3 > c4) && short2 && short1 && (l3 > o4 && l2 > o4 && l1 > o4) && (bull0 && c0 > c4), false)
    EMIT(P_cdlmorningdojistar,                                    // :1416
         lb2 && lbear2 && doji1 && fmax(o1, c1) < c2 && bull0 && c0 > (c2 + (b2 * A.pen[PEN_MORNINGDOJISTAR])), false)
    EMIT(P_cdlmorningstar,                                       // :1454
         lb2 && lbear2 && short1 && fmax(o1, c1) < c2 && bull0 && c0 > (c2 + (b2 * A.pen[PEN_MORNINGSTAR])), false)
    EMIT(P_cdlonneck, false, lb1 && lbear1 && bull0 && o0 < c1 && NEAR(c0, l1))   // :1490
    EMIT(P_cdlpiercing,                                                    // :1519
         lb1 && lbear1 && bull0 && o0 < c1 && c0 > (c1 + (b1 * A.pen[PEN_PIERCING])) && c0 < o1, false)
    EMIT(P_cdlrickshawman,                                     // :1553
         HAS(f0, CF_DOJI | CF_LUS | CF_LDS) && NEAR(us0, ls0), false)
    EMIT(P_cdlrisefall3methods,               // :1581
         lb4 && lbull4 && short3 && short2 && short1 && (h3 < h4 && h2 < h4 && h1 < h4 && l3 > l4 && l2 > l4 && l1 > l4) && lbull0 && c0 > c4,
         lb4 && lbear4 && short3 && short2 && short1 && (h3 < h4 && h2 < h4 && h1 < h4 && l3 > l4 && l2 > l4 && l1 > l4) && lbear0 && c0 < c4)
    EMIT(P_cdlseparatinglines,                                              // :1647
         lb1 && lbear1 && lbull0 && EQUAL(o0, o1), lb1 && lbull1 && lbear0 && EQUAL(o0, o1))
    EMIT(P_cdlshootingstar, false, lb1 && HAS(f0, CF_SHORT | CF_LUS | CF_VSDS) && bull1)          // :1679
    EMIT(P_cdlshortline, HAS(f0, CF_SHORT | CF_SUS | CF_SDS) && bull0, HAS(f0, CF_SHORT | CF_SUS | CF_SDS) && bear0) // :1709
    EMIT(P_cdlspinningtop, HAS(f0, CF_SHORT | CF_USGB | CF_LSGB) && bull0, HAS(f0, CF_SHORT | CF_USGB | CF_LSGB) && bear0)  // :1738
    EMIT(P_cdlstalledpattern, false,                   // :1766
         lb2 && lbull2 && (lbull1 && c1 > c2) && (bull0 && short0 && c0 > c1) && (o0 > o1 && o0 <= c1))
    EMIT(P_cdlsticksandwich,                                    // :1797
         lb2 && lbear2 && lbull1 && o1 > c2 && lbear0 && EQUAL(c0, c2), false)
    EMIT(P_cdltakuri, HAS(f0, CF_DOJI | CF_VLDS | CF_VSUS), false)                                // :1831
    EMIT(P_cdltasukigap,          // :1856
         lb2 && bull2 && bull1 && o1 > c2 && bear0 && o0 > o1 && o0 < c1 && c0 > o2 && c0 < c2,
         lb2 && bear2 && bear1 && o1 < c2 && bull0 && o0 < o1 && o0 > c1 && c0 < o2 && c0 > c2)
    EMIT(P_cdlthrusting, false,                                            // :1894
         lb1 && lbear1 && bull0 && o0 < c1 && c0 > c1 && c0 < (c1 + (b1 * 0.5)))
    EMIT(P_cdltristar,                                             // :1922
         lb2 && doji2 && doji1 && doji0 && ((o1 + c1) / 2.0) < ((o2 + c2) / 2.0) && ((o0 + c0) / 2.0) > ((o1 + c1) / 2.0),
         lb2 && doji2 && doji1 && doji0 && ((o1 + c1) / 2.0) > ((o2 + c2) / 2.0) && ((o0 + c0) / 2.0) < ((o1 + c1) / 2.0))
    EMIT(P_cdlunique3river,                             // :1964
         lb2 && lbear2 && (bear1 && l1 < l2 && c1 > l1) && (o1 < o2 && o1 > c2) && (bull0 && short0 && c0 < c1), false)
    EMIT(P_cdlupsidegap2crows, false,                             // :1997
         lb2 && lbull2 && (bear1 && o1 > c2 && c1 > c2) && (bear0 && o0 > o1 && c0 > c2 && c0 < c1))
    EMIT(P_cdlxsidegap3methods,   // :2027
         lb2 && bull2 && bull1 && o1 > c2 && bear0 && o0 < c1 && o0 > o1 && c0 > o2 && c0 < c2,
         lb2 && bear2 && bear1 && o1 < c2 && bull0 && o0 > c1 && o0 < o1 && c0 < o2 && c0 > c2)
#undef EMIT
#undef HAS
#undef NEAR
#undef EQUAL
    (void)pstride; (void)pbase;
}

// synthetic candle panel for the bench (device-side, counter-based): a random walk of opens with bodies and
// shadows of mixed sizes on a 0.01 price grid, so that patterns do fire
__global__ void candle_synth_kernel(double *o, double *h, double *l, double *c, int n_symbols, int n_bars, int pitch,
                                    unsigned long long seed) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_symbols) return;
    unsigned long long x = seed * 0x9E3779B97F4A7C15ull + (unsigned long long)s * 0xD1B54A32D192ED03ull + 1;
    auto rnd = [&]() {                     // splitmix64 -> [0, 1)
        x += 0x9E3779B97F4A7C15ull;
        unsigned long long z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    };
    double prev = 50.0 + (s % 97);
    for (int t = 0; t < n_bars; ++t) {
        const double g = rnd();
        double op = prev * (1.0 + (g < 0.15 ? 0.02 : (g < 0.3 ? -0.02 : 0.0)));
        const double k = rnd();
        const double body = op * (k < 0.2 ? 0.0 : (k < 0.5 ? 0.01 * rnd() : 0.03 + 0.09 * rnd()));
        double cl = op + (rnd() < 0.5 ? -body : body);
        const double u = rnd(), d = rnd();
        double hi = fmax(op, cl) + op * (u < 0.4 ? 0.0 : (u < 0.7 ? 0.004 : 0.06 * rnd()));
        double lo = fmin(op, cl) - op * (d < 0.4 ? 0.0 : (d < 0.7 ? 0.004 : 0.06 * rnd()));
        op = rint(op * 100.0) / 100.0; cl = rint(cl * 100.0) / 100.0;
        hi = fmax(rint(hi * 100.0) / 100.0, fmax(op, cl));
        lo = fmax(0.01, fmin(rint(lo * 100.0) / 100.0, fmin(op, cl)));
        const size_t at = (size_t)s * pitch + t;
        o[at] = op; h[at] = hi; l[at] = lo; c[at] = cl;
        prev = fmax(5.0, cl);
    }
}

}  // namespace pqb
