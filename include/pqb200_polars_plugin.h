/* pqb200_polars_plugin.h -- the polars expression-plugin ABI of the reference, served by libpqb200.so.
 *
 * The reference reaches its indicators ONLY through polars' plugin loader:
 *   python/polars_quant/talib/overlap.py:13-18 (and every other shim)
 *       register_plugin_function(args=[exprs..., literal params...], plugin_path=<.so>,
 *                                function_name="<lowercase name>", is_elementwise=False)
 *   src/talib/overlap.rs:127  #[polars_expr(output_type=Float64)] pub fn ema(inputs, kwargs)
 *   src/talib/overlap.rs:46   #[polars_expr(output_type_func=bbands_output)] pub fn bbands(...)
 * The #[polars_expr] attribute (pyo3-polars-derive 0.20 / polars-ffi 0.53 `version_0`, Cargo.lock:
 * 1533-1555, 1140-1141 -- third-party, not vendored in the reference tree) generates, per function,
 * the two C symbols declared below, and per library the version / last-error symbols.  polars
 * dlopen()s the library named by `plugin_path` and resolves exactly these names, so pointing the
 * reference's Python shims at libpqb200.so (`_LIB = .../libpqb200.so`) swaps the Rust indicator
 * engine for the B200 one with no other change: same function names, same struct / field names
 * (`bbands`{bb_upper, bb_middle, bb_lower} overlap.rs:30-38, `macd_res`{macd, macd_signal,
 * macd_hist} momentum.rs:239-247), same parameter intake (pickled kwargs `timeperiod` ... as the
 * Rust structs overlap.rs:11-28 / volatility.rs:12-15 / volume.rs:12-16 declare them, or trailing
 * length-1 literal Series as the Python shims send them, momentum.rs:18 `inputs[3].i64()?.get(0)`),
 * same defaults, inputs cast to Float64 first (overlap.rs:48), errors reported through the
 * last-error string with `return_value` left untouched.
 *
 * The struct layouts are the Arrow C Data Interface plus polars-ffi's SeriesExport; they are
 * restated here from the published interface (the crates are absent from /root/reference):
 * verify against `nm -D` of a real plugin and polars-ffi's src/version_0.rs when integrating.
 * Every call runs on the GPU through the single-column entry points of pqb200.h (no CPU path;
 * without a device the call fails with "PQB_ERR_NO_DEVICE ..." in the last-error message).
 * Calls may come concurrently from polars' rayon workers: the shared engine serialises them.
 */
#ifndef PQB200_POLARS_PLUGIN_H
#define PQB200_POLARS_PLUGIN_H

#include <stddef.h>
#include <stdint.h>
#include "pqb200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowSchema {
    const char *format;
    const char *name;
    const char *metadata;
    int64_t flags;
    int64_t n_children;
    struct ArrowSchema **children;
    struct ArrowSchema *dictionary;
    void (*release)(struct ArrowSchema *);
    void *private_data;
};
struct ArrowArray {
    int64_t length;
    int64_t null_count;
    int64_t offset;
    int64_t n_buffers;
    int64_t n_children;
    const void **buffers;
    struct ArrowArray **children;
    struct ArrowArray *dictionary;
    void (*release)(struct ArrowArray *);
    void *private_data;
};
#endif

/* polars-ffi version_0::SeriesExport: one Series = a field + `len` chunk arrays.  The importer takes
 * the chunk arrays (it releases each one itself) and then calls `release`, which frees the
 * containers only. */
typedef struct pqb_series_export {
    struct ArrowSchema *field;
    struct ArrowArray **arrays;
    size_t len;
    void (*release)(struct pqb_series_export *);
    void *private_data;
} pqb_series_export;

/* (major << 16) | minor of the plugin calling convention: 0.1 (return value by pointer + CallerContext) */
PQB_API uint32_t _polars_plugin_get_version(void);
/* message of the last failed call on this thread ("" if none); valid until the thread's next call */
PQB_API const char *_polars_plugin_get_last_error_message(void);

/* One pair per reference function.  `inputs[n_inputs]`: the evaluated input Series (data columns first,
 * then optional length-1 literal parameters in the Python shim's order); `kwargs`: pickle bytes of a
 * dict (may be empty); on success *return_value is filled (caller releases it), on failure it is left
 * untouched and the last-error message is set.  `ctx` (polars' CallerContext) is not used. */
#define PQB_POLARS_PLUGIN(name)                                                                      \
    PQB_API void _polars_plugin_##name(pqb_series_export *inputs, size_t n_inputs, const uint8_t *kwargs, \
                                       size_t kwargs_len, pqb_series_export *return_value, void *ctx);    \
    PQB_API void _polars_plugin_field_##name(struct ArrowSchema *input_fields, size_t n_fields,          \
                                             struct ArrowSchema *return_field);

/* overlap.rs */
PQB_POLARS_PLUGIN(sma)       /* :494  (real; timeperiod=30) */
PQB_POLARS_PLUGIN(ema)       /* :128  (real; timeperiod=30) */
PQB_POLARS_PLUGIN(tema)      /* :513  (real; timeperiod=30) */
PQB_POLARS_PLUGIN(trima)     /* :522  (real; timeperiod=30) */
PQB_POLARS_PLUGIN(ma)        /* :146  (real; timeperiod=30, matype=0) */
PQB_POLARS_PLUGIN(bbands)    /* :47   (real; timeperiod=20, nbdevup=2.0, nbdevdn=2.0) -> struct bbands */
PQB_POLARS_PLUGIN(midpoint)  /* :180  (real; timeperiod=14) */
PQB_POLARS_PLUGIN(midprice)  /* :281  (high, low; timeperiod=14) */
/* momentum.rs */
PQB_POLARS_PLUGIN(rsi)       /* :507  (real; timeperiod=14) */
PQB_POLARS_PLUGIN(macd)      /* :250  (real; fastperiod=12, slowperiod=26, signalperiod=9) -> struct macd_res */
PQB_POLARS_PLUGIN(willr)     /* :630  (high, low, close; timeperiod=14) */
PQB_POLARS_PLUGIN(mom)       /* :384  (real; timeperiod=10) */
PQB_POLARS_PLUGIN(roc)       /* :439  (real; timeperiod=10) */
PQB_POLARS_PLUGIN(rocp)      /* :456 */
PQB_POLARS_PLUGIN(rocr)      /* :473 */
PQB_POLARS_PLUGIN(rocr100)   /* :490 */
PQB_POLARS_PLUGIN(cmo)       /* :181  (real; timeperiod=14) */
PQB_POLARS_PLUGIN(mfi)       /* :286  (high, low, close, volume; timeperiod=14) */
PQB_POLARS_PLUGIN(cci)       /* :138  (high, low, close; timeperiod=14) */
/* volatility.rs */
PQB_POLARS_PLUGIN(trange)    /* :51   (high, low, close) */
PQB_POLARS_PLUGIN(atr)       /* :18   (high, low, close; timeperiod=14) */
PQB_POLARS_PLUGIN(natr)      /* :34   (high, low, close; timeperiod=14) */
/* volume.rs */
PQB_POLARS_PLUGIN(obv)       /* :70   (real, volume) */
PQB_POLARS_PLUGIN(ad)        /* :19   (high, low, close, volume) */
PQB_POLARS_PLUGIN(adosc)     /* :34   (high, low, close, volume; fastperiod=3, slowperiod=10) */
/* Python-level compositions of the reference served as one call (python momentum.py:178-186) */
PQB_POLARS_PLUGIN(stoch)     /* :178 (high, low, close; fastk_period=5, slowk_period=3, slowk_matype=0, slowd_period=3, slowd_matype=0) -> struct stoch{slowk, slowd} */
PQB_POLARS_PLUGIN(stochf)    /* :188 (high, low, close; fastk_period=5, fastd_period=3, fastd_matype=0) -> struct stochf{fastk, fastd} */
PQB_POLARS_PLUGIN(stochrsi)  /* :197 (real; timeperiod=14, fastk_period=5, fastd_period=3, fastd_matype=0) -> struct stochrsi{fastk_rsi, fastd_rsi} */
PQB_POLARS_PLUGIN(macdext)   /* :83  (real; fastperiod=12, fastmatype=0, slowperiod=26, slowmatype=0, signalperiod=9, signalmatype=0) -> struct macdext{macd_dif, macd_dea, macd_hist} */
PQB_POLARS_PLUGIN(kdj)       /* (high, low, close; fastk_period=9, k_period=3, d_period=3) -> struct kdj{k, d, j} */

/* the directional-movement family, momentum.rs; all (…; timeperiod=14) */
PQB_POLARS_PLUGIN(adx)       /* :11   (high, low, close) */
PQB_POLARS_PLUGIN(adxr)      /* :29   (high, low, close) */
PQB_POLARS_PLUGIN(dx)        /* :226  (high, low, close) */
PQB_POLARS_PLUGIN(plus_di)   /* :401  (high, low, close) -- returns calc_dm().0, i.e. DX, like the reference */
PQB_POLARS_PLUGIN(minus_di)  /* :346  (high, low, close) */
PQB_POLARS_PLUGIN(plus_dm)   /* :418  (high, low) */
PQB_POLARS_PLUGIN(minus_dm)  /* :362  (high, low) */
PQB_POLARS_PLUGIN(trix)      /* :544  (real; timeperiod=30) */
PQB_POLARS_PLUGIN(ultosc)    /* :573  (high, low, close; timeperiod1=7, timeperiod2=14, timeperiod3=28) */
PQB_POLARS_PLUGIN(aroon)     /* :63   (high, low; timeperiod=14) -> struct aroon{aroon_up, aroon_down} */

/* candle functions: pattern.rs:9-2065 `#[polars_expr(output_type=Int32)] pub fn cdl*(inputs)`
 * (open, high, low, close [, penetration Float64 literal]) -> Int32 in {-100, 0, 100}, in the reference's order of
 * definition = pattern ids 0..60 of pqb200.h; price.rs:10-91 and momentum.rs:113 (bop) -> Float64 */
#define PQB_CDL_LIST(X)                                                                                      \
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
#define X(n) PQB_POLARS_PLUGIN(n)
PQB_CDL_LIST(X)
#undef X
PQB_POLARS_PLUGIN(avgprice)  /* price.rs:10   (open, high, low, close) */
PQB_POLARS_PLUGIN(medprice)  /* price.rs:35   (high, low) */
PQB_POLARS_PLUGIN(typprice)  /* price.rs:55   (high, low, close) */
PQB_POLARS_PLUGIN(wclprice)  /* price.rs:75   (high, low, close) */
PQB_POLARS_PLUGIN(bop)       /* momentum.rs:113 (open, high, low, close) */

#ifdef __cplusplus
}
#endif
#endif
