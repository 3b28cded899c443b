// windows.cuh -- the rolling max / min suite over MANY windows in one launch (BASELINE config 5: KDJ(k) for k in
// 5 / 9 / 14 / 60 / 250, WILLR / MIDPRICE / Donchian(p) for p in 5 / 20 / 55 / 250, ATR(14) on 10,000 x 5,040).
//
// The fused suite kernel carries ONE period per indicator, so this set used to cost nine launches, each re-reading high /
// low / close and each latency-bound (two or three role warps per CTA).  Here every (indicator, window) pair is a UNIT;
// the host deals the units to G groups, CTA (b, g) runs group g's units (one warp each, lane = symbol, serial walk in the
// reference's operation order: every output bit-identical to the oracle) for symbol block b, fed by its own TMA producer
// warp over the same staged high / low / close bars.  More, smaller CTAs than one per symbol block: the grid (n_blocks x
// G) spreads evenly over the 148 SMs and an SM holds several CTAs, so the serial chains of many windows overlap.
//   KDJ(k, sk, sd)   momentum.py:178-186 + SURVEY D3: polars rolling max / min (full window), fastk, two calc_sma passes
//   WMD(p)           willr momentum.rs:630-662, midprice overlap.rs:281-404, Donchian upper / lower: one van
//                    Herk / Gil-Werman pair of arrays serves all four lines
//   ATR(p)           atr volatility.rs:18-31: calc_trange + calc_ema(trange, 2p - 1)
// van Herk arrays (p + 1 slots of 32 lanes per array) live in shared memory up to W_SMEM_MAX bars; longer windows (250)
// keep them in an L2-resident global scratch (ExtG below: coalesced 256-byte rows, the suffix value of the next bar
// prefetched a bar ahead) so that a 250-bar window costs no shared memory and no occupancy.
#pragma once
#include "suite_kernel.cuh"

namespace pqb {

constexpr int W_MAX_UNITS = 5;        // role warps per CTA (+ 1 producer warp)
constexpr int W_MAX_GROUPS = 4;
constexpr int W_THREADS = 32 * (W_MAX_UNITS + 1);
constexpr int W_SMEM_MAX = 32;        // windows up to this many bars keep their van Herk arrays in shared memory (PQB_WIN_SMEM_MAX)
constexpr int W_FIELDS = 3;           // close, high, low
constexpr int W_STAGE_BYTES = W_FIELDS * SB * SYM * 8;      // 6 KB

enum WinKind { WK_NONE = 0, WK_KDJ = 1, WK_WMD = 2, WK_ATR = 3 };

struct WinUnit {
    int kind, w;                       // window (KDJ fastk_period / WMD timeperiod / ATR timeperiod)
    int sk, sd;                        // KDJ smoothings (calc_sma periods)
    double inv_sk, inv_sd;             // 1.0 / p (overlap.rs:880)
    int ep;                            // ATR: 2p - 1 (volatility.rs:30)
    double alpha;                      // ATR: 2 / (ep + 1)
    int off_h, off_l;                  // shared-memory van Herk arrays (doubles from the ring area) or -1: global scratch
    int off_fk, off_sk;                // KDJ: the two SMA windows
    double *gh, *gl;                   // global scratch of this unit: [n_blocks][w + 1][32]
    double *out[4];                    // KDJ: K, D, J | WMD: willr, midprice, donchian_upper, donchian_lower | ATR: atr
};

struct WinArgs {
    const double *in[W_FIELDS];        // close, high, low (tiled planes)
    const int *start;                  // per-symbol first valid bar
    WinUnit u[W_MAX_GROUPS][W_MAX_UNITS];
    int n_units[W_MAX_GROUPS];
    int smem_bytes[W_MAX_GROUPS];
    int n_groups;
    int steady_lead;                   // a lane is past every warm-up once t - start >= steady_lead
    int n_symbols, n_bars, n_blocks, bars_padded;
};

// van Herk / Gil-Werman in an L2-resident global scratch: same algorithm as Ext (suite_kernel.cuh), rows of 32 lanes.
// Block end (once per p bars): raw values -> suffix extremes in place, newest to oldest; loads batched 8 ahead of the max
// chain.  Out of line: the steady loop keeps only the per-bar work (and its registers).
__device__ __noinline__ void extg_rebuild(double *hb, double *lb, int p) {
    double sh = vmin(), sl = vmax();
    int q = p;
    while (q > 0) {
        const int nb = min(q, 8);
        double a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < nb) { a[i] = __ldcg(hb + (q - 1 - i) * SYM); b[i] = __ldcg(lb + (q - 1 - i) * SYM); }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < nb) {
                sh = dmax(sh, a[i]);
                sl = dmin(sl, b[i]);
                __stcg(hb + (q - 1 - i) * SYM, sh);
                __stcg(lb + (q - 1 - i) * SYM, sl);
            }
        q -= nb;
    }
}
struct ExtG {
    double *hb, *lb;
    int off, p;
    double ph, pl;                     // running prefix extremes of the current block
    double nh, nl, nh2, nl2;           // the previous block's suffix extremes at slots off + 1 and off + 2: loaded TWO bars ahead
                                       // of their use, so that an L2 round trip (~250 cycles + queueing) hides behind two bars of arithmetic
    __device__ __forceinline__ void init(double *h, double *l, int p_, int lane) {
        hb = h + lane;
        lb = l + lane;
        p = p_;
        off = 0;
        ph = vmin();
        pl = vmax();
        for (int q = 0; q <= p; ++q) {
            __stcg(hb + q * SYM, vmin());
            __stcg(lb + q * SYM, vmax());
        }
        nh = nh2 = vmin();
        nl = nl2 = vmax();
    }
    __device__ __forceinline__ void step(double h, double l, double &hn, double &ln) {
        ph = dmax(ph, h);
        pl = dmin(pl, l);
        hn = dmax(ph, nh);
        ln = dmin(pl, nl);
        __stcg(hb + off * SYM, h);
        __stcg(lb + off * SYM, l);
        ++off;
        if (off == p) {
            extg_rebuild(hb, lb, p);
            ph = vmin();
            pl = vmax();
            off = 0;
            nh2 = __ldcg(hb + min(1, p) * SYM);              // (slot p is the sentinel)
            nl2 = __ldcg(lb + min(1, p) * SYM);
        }
        nh = nh2;
        nl = nl2;
        const int q = min(off + 2, p);
        nh2 = __ldcg(hb + q * SYM);
        nl2 = __ldcg(lb + q * SYM);
    }
};

struct WCtx {
    size_t pos;        // element offset of (this lane, current bar)
    int a, n_bars;     // this lane's first valid bar; bars of the panel
};

// ---- KDJ(k, sk, sd) ---------------------------------------------------------------------------------------------------------
template <class EXT>
struct UnitKdj {
    EXT ek;
    Ring fr, sr;
    double s_k, s_d;
    struct { int w, sk, sd; double inv_sk, inv_sd; double *out[3]; } U;      // this unit's parameters, in registers
    __device__ __forceinline__ void take(const WinUnit &u) {
        U.w = u.w; U.sk = u.sk; U.sd = u.sd; U.inv_sk = u.inv_sk; U.inv_sd = u.inv_sd;
        U.out[0] = u.out[0]; U.out[1] = u.out[1]; U.out[2] = u.out[2];
    }
    template <bool STEADY>
    __device__ __forceinline__ void step(const WCtx &X, int t, double c, double h, double l) {
        const int j = t - X.a;
        const bool live = STEADY || t < X.n_bars;
        const bool in = STEADY || (j >= 0 && live);
        const double nn = qnan();
        double hn, ln;
        ek.step(in ? h : ninf(), in ? l : pinf(), hn, ln);
        double ok_ = nn, od = nn, oj = nn;
        const int j1 = j - (U.w - 1);                     // index in the fastk series (polars rolling: k-1 nulls)
        const bool v1 = STEADY || (j1 >= 0 && live);
        // momentum.py:183 -- IEEE x / 0 (= x * inf: +-inf, or NaN for 0 / 0) without the slow path
        const double num = (c - ln) * 100.0, den = hn - ln;
        const double fk = (den == 0.0) ? num * copysign(pinf(), den) : num / den;
        const double oldf = fr.swap(fk);
        double sk = 0.0;
        const int j2 = j1 - (U.sk - 1);
        if (v1) {
            s_k += fk;                                    // slowk = calc_sma(fastk, sk) overlap.rs:871
            if (STEADY || j1 >= U.sk) s_k -= oldf;
            sk = s_k * U.inv_sk;
        }
        const double olds = sr.swap(sk);
        if (v1 && (STEADY || j2 >= 0)) {
            ok_ = sk;
            s_d += sk;                                    // slowd = calc_sma(slowk, sd)
            if (STEADY || j2 >= U.sd) s_d -= olds;
            if (STEADY || j2 >= U.sd - 1) {
                const double sd = s_d * U.inv_sd;
                od = sd;
                oj = 3.0 * sk - 2.0 * sd;                 // J = 3K - 2D (D3)
            }
        }
        stg(U.out[0] + X.pos, ok_);
        stg(U.out[1] + X.pos, od);
        stg(U.out[2] + X.pos, oj);
    }
};

// ---- WILLR / MIDPRICE / Donchian(p) -----------------------------------------------------------------------------------------
template <class EXT>
struct UnitWmd {
    EXT ew;
    struct { int w; double *out[4]; } U;
    __device__ __forceinline__ void take(const WinUnit &u) {
        U.w = u.w;
        for (int q = 0; q < 4; ++q) U.out[q] = u.out[q];
    }
    template <bool STEADY>
    __device__ __forceinline__ void step(const WCtx &X, int t, double c, double h, double l) {
        const int j = t - X.a;
        const bool live = STEADY || t < X.n_bars;
        const bool in = STEADY || (j >= 0 && live);
        const double nn = qnan();
        double hn, ln;
        ew.step(in ? h : ninf(), in ? l : pinf(), hn, ln);
        if (U.out[0]) {                                   // willr momentum.rs:630-662
            double o = nn;
            if ((STEADY || j >= U.w - 1) && live) {
                const double diff = hn - ln;
                const bool z = diff == 0.0;
                const double q = -100.0 * (hn - c) / (z ? 1.0 : diff);                    // :653-657
                o = z ? 0.0 : q;
            }
            stg(U.out[0] + X.pos, o);
        }
        if (U.out[1]) stg(U.out[1] + X.pos, in ? (hn + ln) / 2.0 : nn);                   // midprice overlap.rs:401
        if (U.out[2]) stg(U.out[2] + X.pos, in ? hn : nn);                                // Donchian upper / lower (D3)
        if (U.out[3]) stg(U.out[3] + X.pos, in ? ln : nn);
    }
};

// ---- ATR(p) -------------------------------------------------------------------------------------------------------------------
struct UnitAtr {
    Ema atr;
    double pc;
    struct { int ep; double alpha; double *out[1]; } U;
    __device__ __forceinline__ void take(const WinUnit &u) { U.ep = u.ep; U.alpha = u.alpha; U.out[0] = u.out[0]; }
    template <bool STEADY>
    __device__ __forceinline__ void step(const WCtx &X, int t, double c, double h, double l) {
        const int j = t - X.a;
        const bool live = STEADY || t < X.n_bars;
        const double tr = rs_max(rs_max(h - l, fabs(h - pc)), fabs(l - pc));             // volatility.rs:77
        const bool ok = atr.step<STEADY>(tr, j - 1, U.ep, U.alpha);                      // :30 calc_ema(trange, 2p-1)
        stg(U.out[0] + X.pos, (ok && live) ? atr.y : qnan());
        pc = c;
    }
};

template <class UNIT>
__device__ __forceinline__ void run_unit(UNIT &R, const WinArgs &A, const WinUnit &U_, uint32_t full, uint32_t empty,
                                         double *ring_smem, int block, int lane) {
    R.take(U_);
    const int sym = block * SYM + lane;
    const int a = A.start ? A.start[(sym < A.n_symbols) ? sym : block * SYM] : 0;
    const int src_lane = (sym < A.n_symbols) ? lane : 0;   // lanes past the last symbol follow lane 0 (no slow-path divisions)
    WCtx X{(size_t)block * A.bars_padded * SYM + lane, a, A.n_bars};
    int amax = a;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) amax = max(amax, __shfl_xor_sync(FULL, amax, d));
    const long long steady_from = (long long)amax + A.steady_lead;
    const int n_iter = A.bars_padded / SB;
    for (int it = 0; it < n_iter; ++it) {
        const int st = it % NS;
        mbar_wait(full + st * 8, (it / NS) & 1);
        const uint32_t sp = st * W_STAGE_BYTES + src_lane * 8;
        const int t0 = it * SB;
        if (t0 >= steady_from && t0 + SB <= A.n_bars) {
#pragma unroll 1
            for (int b = 0; b < SB; ++b) {
                const uint32_t q = sp + b * (SYM * 8);
                R.template step<true>(X, t0 + b, lds(q), lds(q + 1 * SB * SYM * 8), lds(q + 2 * SB * SYM * 8));
                X.pos += SYM;
            }
        } else {
#pragma unroll 1
            for (int b = 0; b < SB; ++b) {
                if (t0 + b < A.n_bars) {
                    const uint32_t q = sp + b * (SYM * 8);
                    R.template step<false>(X, t0 + b, lds(q), lds(q + 1 * SB * SYM * 8), lds(q + 2 * SB * SYM * 8));
                }
                X.pos += SYM;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + st * 8);
    }
}

// grid = n_blocks * n_groups; CTA (b, g): warps 0..n_units-1 = units of group g, warp W_MAX_UNITS = TMA producer
__global__ void __launch_bounds__(W_THREADS, 3) window_suite_kernel(const __grid_constant__ WinArgs A) {
    uint64_t *full_p = reinterpret_cast<uint64_t *>(smem_dyn + NS * W_STAGE_BYTES);
    uint64_t *empty_p = full_p + NS;
    double *rings = reinterpret_cast<double *>(empty_p + NS);
    const uint32_t stage = smem_u32(smem_dyn), full = smem_u32(full_p), empty = smem_u32(empty_p);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // group-major: the CTAs of the heaviest group (group 0) are scheduled first, the lighter groups fill in behind them
    const int g = (int)(blockIdx.x / (unsigned)A.n_blocks), block = (int)(blockIdx.x % (unsigned)A.n_blocks);
    const int n_units = A.n_units[g];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full_p[s], 1);
            mbar_init(&empty_p[s], n_units);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == W_MAX_UNITS) {
        if (lane == 0) {
            const int n_iter = A.bars_padded / SB;
            const size_t base = (size_t)block * A.bars_padded * SYM;
            for (int it = 0; it < n_iter; ++it) {
                const int st = it % NS;
                if (it >= NS) mbar_wait(empty + st * 8, ((it / NS) & 1) ^ 1);
                mbar_expect_tx(full + st * 8, (uint32_t)W_STAGE_BYTES);
                const size_t off = base + (size_t)it * SB * SYM;
#pragma unroll
                for (int f = 0; f < W_FIELDS; ++f)
                    tma_load_1d(stage + st * W_STAGE_BYTES + f * SB * SYM * 8, A.in[f] + off, (uint32_t)(SB * SYM * sizeof(double)),
                                full + st * 8);
            }
        }
        return;
    }
    if (warp >= n_units) return;
    const WinUnit &U = A.u[g][warp];
    const size_t gbase = (size_t)block * (size_t)(U.w + 1) * SYM;
    if (U.kind == WK_KDJ) {
        if (U.off_h >= 0) {
            UnitKdj<Ext> R;
            R.ek.init(rings + U.off_h, rings + U.off_l, U.w, lane);
            R.fr.init(rings + U.off_fk, U.sk, lane);
            R.sr.init(rings + U.off_sk, U.sd, lane);
            R.s_k = R.s_d = 0.0;
            __syncwarp();
            run_unit(R, A, U, full, empty, rings, block, lane);
        } else {
            UnitKdj<ExtG> R;
            R.ek.init(U.gh + gbase, U.gl + gbase, U.w, lane);
            R.fr.init(rings + U.off_fk, U.sk, lane);
            R.sr.init(rings + U.off_sk, U.sd, lane);
            R.s_k = R.s_d = 0.0;
            __syncwarp();
            run_unit(R, A, U, full, empty, rings, block, lane);
        }
    } else if (U.kind == WK_WMD) {
        if (U.off_h >= 0) {
            UnitWmd<Ext> R;
            R.ew.init(rings + U.off_h, rings + U.off_l, U.w, lane);
            __syncwarp();
            run_unit(R, A, U, full, empty, rings, block, lane);
        } else {
            UnitWmd<ExtG> R;
            R.ew.init(U.gh + gbase, U.gl + gbase, U.w, lane);
            __syncwarp();
            run_unit(R, A, U, full, empty, rings, block, lane);
        }
    } else {
        UnitAtr R;
        R.atr.init();
        R.pc = 0.0;
        run_unit(R, A, U, full, empty, rings, block, lane);
    }
}

}  // namespace pqb
