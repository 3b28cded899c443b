/*
 * pqb200.h -- C ABI of the B200-native polars-quant indicator engine (libpqb200.so).
 *
 * This is the drop-in boundary for ONE path of Firstastor/polars-quant: the src/talib
 * indicator engine applied to wide `{symbol}_{column}` f64 panels.  Everything is plain C:
 * pointers, sizes, int status codes; no C++/torch types cross it and nothing throws.
 *
 * Two layers are exported:
 *   (A) engine ABI `pqb_*`  -- what a host (the unchanged Rust crate through a thin
 *       `extern "C"` block, or ctypes) binds to run a whole panel in one fused launch.
 *       It replaces, per symbol, the chain of per-column plugin calls the reference makes
 *       (SURVEY.md 3.3): sma overlap.rs:494, ema :128, tema :513, trima :522, bbands :47,
 *       macd momentum.rs:250, rsi :507, trange volatility.rs:51, atr :18, natr :34,
 *       obv volume.rs:70, ad :19, STOCH/KDJ momentum.py:178-186, willr momentum.rs:630,
 *       midprice overlap.rs:281.
 *   (B) polars-plugin symbols `_polars_plugin_<name>` / `_polars_plugin_field_<name>` /
 *       `_polars_plugin_get_version` / `_polars_plugin_get_last_error_message`, with the
 *       same names the reference's `#[polars_expr]` macro emits (pyo3-polars 0.26), so the
 *       reference's Python shims (python/polars_quant/talib/ *.py, `register_plugin_function`)
 *       can point `_LIB` at libpqb200.so unchanged.  Declared in pqb200_plugin.h.
 *
 * There is NO CPU fallback: every compute entry point fails with PQB_ERR_NO_DEVICE when no
 * sm_100 device / driver is present.
 *
 * Column conventions (identical to the reference): all data f64; outputs have
 * the input's length; warm-up positions are Arrow NULLs (validity bit 0; the value slot
 * holds a quiet NaN); validity bitmaps are Arrow LSB-first bitmaps.
 */
#ifndef PQB200_H
#define PQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQB_ABI_VERSION 1
#if defined(__GNUC__)
#define PQB_API __attribute__((visibility("default")))
#else
#define PQB_API
#endif

/* ---- status codes ------------------------------------------------------------------- */
enum {
    PQB_OK = 0,
    PQB_ERR_NO_DEVICE = -1,   /* no CUDA driver / no sm_100 device: the engine never falls back */
    PQB_ERR_CUDA = -2,        /* a CUDA call failed; text in pqb_last_error() */
    PQB_ERR_INVALID = -3,     /* bad argument (null pointer, negative size, unknown id ...) */
    PQB_ERR_UNSUPPORTED = -4, /* valid in the reference but outside this build (e.g. window > limit) */
    PQB_ERR_NULLS = -5,       /* input has nulls where the reference returns Err (cont_slice()?) */
    PQB_ERR_ALLOC = -6
};

/* ---- panel fields (inputs) ------------------------------------------------------------ */
enum pqb_field { PQB_CLOSE = 0, PQB_HIGH = 1, PQB_LOW = 2, PQB_VOLUME = 3, PQB_N_FIELDS = 4 };

/* ---- suite outputs: 15 indicators, 21 f64 columns ---------------------- */
enum pqb_output {
    PQB_OUT_SMA = 0,        /* calc_sma   overlap.rs:871  */
    PQB_OUT_EMA = 1,        /* calc_ema   overlap.rs:660  */
    PQB_OUT_TEMA = 2,       /* calc_tema  overlap.rs:1177 */
    PQB_OUT_TRIMA = 3,      /* calc_trima overlap.rs:1313 */
    PQB_OUT_BB_UPPER = 4,   /* bbands     overlap.rs:47 (struct field bb_upper) */
    PQB_OUT_BB_MIDDLE = 5,
    PQB_OUT_BB_LOWER = 6,
    PQB_OUT_MACD = 7,       /* macd       momentum.rs:250 (struct field macd) */
    PQB_OUT_MACD_SIGNAL = 8,
    PQB_OUT_MACD_HIST = 9,
    PQB_OUT_RSI = 10,       /* rsi        momentum.rs:507 */
    PQB_OUT_TRANGE = 11,    /* trange     volatility.rs:51 */
    PQB_OUT_ATR = 12,       /* atr        volatility.rs:18 */
    PQB_OUT_NATR = 13,      /* natr       volatility.rs:34 */
    PQB_OUT_OBV = 14,       /* obv        volume.rs:70 */
    PQB_OUT_AD = 15,        /* ad         volume.rs:19 */
    PQB_OUT_KDJ_K = 16,     /* STOCH slowk momentum.py:178-186 (KDJ K) */
    PQB_OUT_KDJ_D = 17,     /* STOCH slowd */
    PQB_OUT_KDJ_J = 18,     /* 3K - 2D */
    PQB_OUT_WILLR = 19,     /* willr      momentum.rs:630 */
    PQB_OUT_MIDPRICE = 20,  /* midprice   overlap.rs:281 (Donchian mid) */
    PQB_N_OUTPUTS = 21
};

/* Indicator groups: bit i of pqb_suite_params.indicators enables group i. */
enum pqb_indicator {
    PQB_IND_SMA = 1u << 0, PQB_IND_EMA = 1u << 1, PQB_IND_TEMA = 1u << 2, PQB_IND_TRIMA = 1u << 3,
    PQB_IND_BBANDS = 1u << 4, PQB_IND_MACD = 1u << 5, PQB_IND_RSI = 1u << 6, PQB_IND_TRANGE = 1u << 7,
    PQB_IND_ATR = 1u << 8, PQB_IND_NATR = 1u << 9, PQB_IND_OBV = 1u << 10, PQB_IND_AD = 1u << 11,
    PQB_IND_KDJ = 1u << 12, PQB_IND_WILLR = 1u << 13, PQB_IND_MIDPRICE = 1u << 14,
    PQB_IND_ALL = (1u << 15) - 1
};

/* Periods; defaults are the reference's Python signature defaults
 * (python/polars_quant/talib/overlap.py, momentum.py, volatility.py). */
typedef struct pqb_suite_params {
    uint32_t indicators;      /* PQB_IND_* mask */
    int32_t sma_period;       /* 30 */
    int32_t ema_period;       /* 30 */
    int32_t tema_period;      /* 30 */
    int32_t trima_period;     /* 30 */
    int32_t bbands_period;    /* 20 */
    double  bbands_nbdevup;   /* 2.0 */
    double  bbands_nbdevdn;   /* 2.0 */
    int32_t macd_fast;        /* 12 */
    int32_t macd_slow;        /* 26 */
    int32_t macd_signal;      /* 9 */
    int32_t rsi_period;       /* 14 */
    int32_t atr_period;       /* 14 */
    int32_t natr_period;      /* 14 */
    int32_t kdj_fastk;        /* 9  (STOCH fastk_period) */
    int32_t kdj_slowk;        /* 3  (STOCH slowk_period, matype 0) */
    int32_t kdj_slowd;        /* 3  (STOCH slowd_period, matype 0) */
    int32_t willr_period;     /* 14 */
    int32_t midprice_period;  /* 14 */
} pqb_suite_params;

PQB_API void pqb_suite_params_default(pqb_suite_params *p);

/* ---- engine ---------------------------------------------------------------------------- */
typedef struct pqb_engine pqb_engine;   /* one per GPU: device, streams, scratch */
typedef struct pqb_panel pqb_panel;     /* one packed device-resident panel + its outputs */

PQB_API int pqb_abi_version(void);
/* Thread-local text of the last failure on the calling thread ("" if none). */
PQB_API const char *pqb_last_error(void);
/* Number of usable sm_100 devices (0 without driver/GPU; never fails). */
PQB_API int pqb_device_count(void);

PQB_API int pqb_engine_create(int device, pqb_engine **out);
PQB_API void pqb_engine_destroy(pqb_engine *e);

/* ---- panel ----------------------------------------------------------------------------
 * Host side (pinned staging, everything that crosses this ABI): row-major [symbol][pitch] f64,
 * pitch = n_bars rounded up to 16, plus Arrow LSB-first validity bitmaps per output column.
 * Device side: every plane is TILED [symbol block of 32][bar][32 symbols]:
 * element (s, t) lives at ((s/32) * bars_padded + t) * 32 + s%32 doubles, `bars_padded` from
 * pqb_panel_tiled_shape().  Upload / download / pqb_suite_run_host convert on the device.
 * Allocates device planes for the fields in `fields_mask` (bit f = enum pqb_field f), the
 * output planes for `outputs_mask` (bit k = enum pqb_output k), output validity bitmaps and
 * (if `host_staging` != 0) pinned host staging plus the device transfer buffers. */
PQB_API int pqb_panel_create(pqb_engine *e, int64_t n_symbols, int64_t n_bars, uint32_t fields_mask,
                     uint32_t outputs_mask, int host_staging, pqb_panel **out);
PQB_API void pqb_panel_destroy(pqb_panel *p);
PQB_API int64_t pqb_panel_pitch(const pqb_panel *p);          /* doubles per symbol row */
PQB_API int64_t pqb_panel_validity_pitch(const pqb_panel *p); /* bytes per symbol row of a bitmap */
/* Geometry of the tiled device planes: symbol blocks and (padded) bars per block. */
PQB_API int pqb_panel_tiled_shape(const pqb_panel *p, int64_t *n_blocks, int64_t *bars_padded);

/* Copies one Arrow column (values + optional validity bitmap + bit/element offset, as in the
 * Arrow C Data Interface the reference's plugin boundary receives) into the
 * pinned staging row of (`symbol`, `field`).  Leading/trailing nulls set the symbol's valid
 * range; interior nulls -> PQB_ERR_UNSUPPORTED (see DESIGN.md "nulls").  */
PQB_API int pqb_panel_set_column(pqb_panel *p, int64_t symbol, int field, const double *values,
                         const uint8_t *validity, int64_t offset, int64_t len);
/* Direct access to the pinned staging planes so a loader can write `[symbol][pitch]` rows in
 * place (zero-copy load()); NULL if the panel has no host staging. */
PQB_API double *pqb_panel_host_field(pqb_panel *p, int field);
PQB_API const double *pqb_panel_host_output(pqb_panel *p, int output);
PQB_API const uint8_t *pqb_panel_host_validity(pqb_panel *p, int output);
/* Per-symbol first valid bar (leading nulls), default 0 for every symbol. */
PQB_API int pqb_panel_set_starts(pqb_panel *p, const int32_t *starts /* [n_symbols] */);

/* Transfer path (async on the engine's stream; pqb_panel_sync to wait). */
PQB_API int pqb_panel_upload(pqb_panel *p);      /* pinned staging -> device, all fields */
PQB_API int pqb_panel_download(pqb_panel *p);    /* device outputs + validity -> pinned staging */
PQB_API int pqb_panel_sync(pqb_panel *p);

/* The hot path: ONE fused pass over close/high/low/volume per symbol computing every enabled
 * indicator (+ one small launch for the validity bitmaps).  Device-resident in, device-resident
 * out; asynchronous on the engine's stream. */
PQB_API int pqb_suite_run(pqb_panel *p, const pqb_suite_params *params);

/* End-to-end convenience = the call a panel-level host makes: chunked, double-buffered
 * upload -> suite -> download over the pinned staging, overlapping both DMA directions with
 * compute; returns after everything landed in host staging. */
PQB_API int pqb_suite_run_host(pqb_panel *p, const pqb_suite_params *params, int64_t chunk_symbols);

/* Copies one output column back into caller buffers (Arrow layout: values[len], validity
 * bitmap of ceil(len/8) bytes, may be NULL).  Requires a prior download + sync. */
PQB_API int pqb_panel_get_output(pqb_panel *p, int64_t symbol, int output, double *values,
                         uint8_t *validity, int64_t len);

/* Device pointers (TILED planes; validity bitmaps are row-major [symbol][validity_pitch]) for
 * callers that stay on the GPU. */
PQB_API const double *pqb_panel_device_field(const pqb_panel *p, int field);
PQB_API const double *pqb_panel_device_output(const pqb_panel *p, int output);
PQB_API const uint8_t *pqb_panel_device_validity(const pqb_panel *p, int output);

/* ---- single-column entry points ---------------------------------------------------------
 * One call == one reference plugin call on one column (BASELINE config 1).  Host buffers in,
 * host buffers out; values[len] + Arrow validity bitmaps (may be NULL on input = no nulls;
 * outputs' validity must be non-NULL, ceil(len/8) bytes).  They run the same fused kernel on a
 * 1-symbol panel with one indicator enabled.  Null handling follows the reference function:
 * momentum.rs functions return PQB_ERR_NULLS on any null (cont_slice()?). */
typedef struct pqb_col {        /* one Arrow f64 column (borrowed) */
    const double *values;
    const uint8_t *validity;    /* LSB-first bitmap or NULL */
    int64_t offset;             /* element offset into values/validity */
    int64_t len;
} pqb_col;
typedef struct pqb_out_col {    /* caller-allocated output column */
    double *values;             /* [len] */
    uint8_t *validity;          /* [ceil(len/8)] */
} pqb_out_col;

PQB_API int pqb_sma(pqb_engine *e, const pqb_col *real, int32_t timeperiod, pqb_out_col *out);
PQB_API int pqb_ema(pqb_engine *e, const pqb_col *real, int32_t timeperiod, pqb_out_col *out);
PQB_API int pqb_tema(pqb_engine *e, const pqb_col *real, int32_t timeperiod, pqb_out_col *out);
PQB_API int pqb_trima(pqb_engine *e, const pqb_col *real, int32_t timeperiod, pqb_out_col *out);
/* matype as calc_ma overlap.rs:857: 0/7/other SMA, 1 EMA, 4 TEMA, 5 TRIMA; 2,3,6,8 unsupported */
PQB_API int pqb_ma(pqb_engine *e, const pqb_col *real, int32_t timeperiod, int32_t matype, pqb_out_col *out);
PQB_API int pqb_bbands(pqb_engine *e, const pqb_col *real, int32_t timeperiod, double nbdevup, double nbdevdn,
               pqb_out_col *upper, pqb_out_col *middle, pqb_out_col *lower);
PQB_API int pqb_macd(pqb_engine *e, const pqb_col *real, int32_t fastperiod, int32_t slowperiod,
             int32_t signalperiod, pqb_out_col *macd, pqb_out_col *signal, pqb_out_col *hist);
PQB_API int pqb_rsi(pqb_engine *e, const pqb_col *real, int32_t timeperiod, pqb_out_col *out);
PQB_API int pqb_trange(pqb_engine *e, const pqb_col *high, const pqb_col *low, const pqb_col *close, pqb_out_col *out);
PQB_API int pqb_atr(pqb_engine *e, const pqb_col *high, const pqb_col *low, const pqb_col *close,
            int32_t timeperiod, pqb_out_col *out);
PQB_API int pqb_natr(pqb_engine *e, const pqb_col *high, const pqb_col *low, const pqb_col *close,
             int32_t timeperiod, pqb_out_col *out);
PQB_API int pqb_obv(pqb_engine *e, const pqb_col *close, const pqb_col *volume, pqb_out_col *out);
PQB_API int pqb_ad(pqb_engine *e, const pqb_col *high, const pqb_col *low, const pqb_col *close,
           const pqb_col *volume, pqb_out_col *out);
/* STOCH with matype 0 for both smoothings (momentum.py:178-186); KDJ adds J = 3K-2D (D3). */
PQB_API int pqb_stoch(pqb_engine *e, const pqb_col *high, const pqb_col *low, const pqb_col *close,
              int32_t fastk_period, int32_t slowk_period, int32_t slowd_period,
              pqb_out_col *slowk, pqb_out_col *slowd);
PQB_API int pqb_kdj(pqb_engine *e, const pqb_col *high, const pqb_col *low, const pqb_col *close,
            int32_t fastk_period, int32_t k_period, int32_t d_period,
            pqb_out_col *k, pqb_out_col *d, pqb_out_col *j);
PQB_API int pqb_willr(pqb_engine *e, const pqb_col *high, const pqb_col *low, const pqb_col *close,
              int32_t timeperiod, pqb_out_col *out);
PQB_API int pqb_midprice(pqb_engine *e, const pqb_col *high, const pqb_col *low, int32_t timeperiod, pqb_out_col *out);

/* ---- measurement helpers (used by bench.py / tests; not needed by a host) -------------- */
/* Fills the device panel with the synthetic random-walk OHLCV of SURVEY.md 8d (and mirrors it
 * into host staging if present and `to_host` != 0). */
PQB_API int pqb_panel_fill_synthetic(pqb_panel *p, uint64_t seed, double sigma, int to_host);
/* Times `iters` back-to-back pqb_suite_run calls with CUDA events on the engine's stream
 * (after `warmup` untimed ones); writes total milliseconds of the whole step and of the
 * dominant fused kernel alone; returns kernels launched per step in *launches_per_step. */
PQB_API int pqb_suite_time(pqb_panel *p, const pqb_suite_params *params, int warmup, int iters,
                   float *ms_total, float *ms_fused_kernel, int *launches_per_step);
/* Same for the end-to-end host path (pqb_suite_run_host); ms includes H2D + kernels + D2H. */
PQB_API int pqb_suite_time_host(pqb_panel *p, const pqb_suite_params *params, int64_t chunk_symbols,
                        int warmup, int iters, float *ms_total);
/* Kernels launched by the most recent pqb_suite_run / pqb_suite_run_host on this panel. */
PQB_API int pqb_panel_last_launches(const pqb_panel *p);
/* Writes > L2-size bytes to a scratch buffer (L2 flush between timed iterations). */
PQB_API int pqb_flush_l2(pqb_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* PQB200_H */
