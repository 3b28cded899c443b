// polars_plugin.cu -- the reference's polars expression-plugin ABI (include/pqb200_polars_plugin.h),
// host code only.  Each `_polars_plugin_<name>` imports its input Series through the Arrow C Data
// Interface (any number of chunks, offsets, validity bitmaps; numeric dtypes are cast to Float64 like
// `inputs[0].cast(&DataType::Float64)?`, overlap.rs:48), takes the parameters from pickled kwargs or
// trailing length-1 literal Series, runs the matching single-column entry point of pqb200.h on the GPU
// and exports the result (a Float64 array, or the reference's struct of Float64 fields).  No CPU path.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pqb200_polars_plugin.h"

namespace {

thread_local std::string g_perr;

void set_err(const char *fmt, ...) {
    char buf[768];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_perr = buf;
}

// ---- the shared engine (polars may call from several rayon workers; pqb single-column calls lock it) ----
std::mutex g_engine_mu;
pqb_engine *g_engine = nullptr;

pqb_engine *engine() {
    std::lock_guard<std::mutex> lk(g_engine_mu);
    if (!g_engine) {
        int dev = 0;
        if (const char *s = getenv("PQB_DEVICE")) dev = atoi(s);
        pqb_engine *e = nullptr;
        if (pqb_engine_create(dev, &e) != PQB_OK) {
            set_err("%s", pqb_last_error());
            return nullptr;
        }
        g_engine = e;
    }
    return g_engine;
}

// ---- inputs ---------------------------------------------------------------------------------------
inline bool bit_at(const uint8_t *bm, int64_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }

double value_at(const char fmt, const void *buf, int64_t i) {
    switch (fmt) {
        case 'g': return static_cast<const double *>(buf)[i];
        case 'f': return static_cast<const float *>(buf)[i];
        case 'l': return (double)static_cast<const int64_t *>(buf)[i];
        case 'L': return (double)static_cast<const uint64_t *>(buf)[i];
        case 'i': return (double)static_cast<const int32_t *>(buf)[i];
        case 'I': return (double)static_cast<const uint32_t *>(buf)[i];
        case 's': return (double)static_cast<const int16_t *>(buf)[i];
        case 'S': return (double)static_cast<const uint16_t *>(buf)[i];
        case 'c': return (double)static_cast<const int8_t *>(buf)[i];
        case 'C': return (double)static_cast<const uint8_t *>(buf)[i];
    }
    return 0.0;
}

struct Column {                       // one input Series as a pqb_col (borrowed or materialised)
    std::vector<double> vals;
    std::vector<uint8_t> valid;
    pqb_col col{nullptr, nullptr, 0, 0};
    std::string name;
};

bool numeric_format(const char *f) { return f && f[0] && !f[1] && strchr("gflLiIsScC", f[0]); }

int64_t series_len(const pqb_series_export *s) {
    int64_t n = 0;
    for (size_t k = 0; k < s->len; ++k) n += s->arrays[k]->length;
    return n;
}

bool import_column(const pqb_series_export *s, int idx, Column *out) {
    if (!s || !s->field || (!s->arrays && s->len)) { set_err("input %d: malformed SeriesExport", idx); return false; }
    const char *fmt = s->field->format;
    if (!numeric_format(fmt)) {
        set_err("input %d (%s): dtype with Arrow format '%s' cannot be cast to Float64", idx,
                s->field->name ? s->field->name : "", fmt ? fmt : "?");
        return false;
    }
    out->name = s->field->name ? s->field->name : "";
    const int64_t n = series_len(s);
    for (size_t k = 0; k < s->len; ++k) {
        const ArrowArray *a = s->arrays[k];
        if (a->n_buffers != 2 || (a->length > 0 && !a->buffers[1])) { set_err("input %d: not a primitive array", idx); return false; }
    }
    if (s->len == 1 && fmt[0] == 'g') {                       // zero-copy: one Float64 chunk
        const ArrowArray *a = s->arrays[0];
        out->col.values = static_cast<const double *>(a->buffers[1]);
        out->col.validity = (a->null_count != 0) ? static_cast<const uint8_t *>(a->buffers[0]) : nullptr;
        out->col.offset = a->offset;
        out->col.len = n;
        return true;
    }
    out->vals.resize((size_t)std::max<int64_t>(n, 1));
    bool any_valid = false;
    for (size_t k = 0; k < s->len; ++k) any_valid |= (s->arrays[k]->null_count != 0 && s->arrays[k]->buffers[0]);
    if (any_valid) out->valid.assign((size_t)((n + 7) / 8), 0xff);
    int64_t pos = 0;
    for (size_t k = 0; k < s->len; ++k) {                     // state carries across chunks (overlap.rs:674)
        const ArrowArray *a = s->arrays[k];
        const uint8_t *bm = (a->null_count != 0) ? static_cast<const uint8_t *>(a->buffers[0]) : nullptr;
        for (int64_t i = 0; i < a->length; ++i, ++pos) {
            out->vals[(size_t)pos] = value_at(fmt[0], a->buffers[1], a->offset + i);
            if (bm && !bit_at(bm, a->offset + i)) out->valid[(size_t)(pos >> 3)] &= (uint8_t)~(1u << (pos & 7));
        }
    }
    out->col.values = out->vals.data();
    out->col.validity = any_valid ? out->valid.data() : nullptr;
    out->col.offset = 0;
    out->col.len = n;
    return true;
}

// a trailing literal parameter: first value of a numeric Series, null / empty -> not given
bool literal_value(const pqb_series_export *s, double *v) {
    if (!s || !s->field || !numeric_format(s->field->format)) return false;
    for (size_t k = 0; k < s->len; ++k) {
        const ArrowArray *a = s->arrays[k];
        if (a->length == 0) continue;
        if (a->null_count != 0 && a->buffers[0] && !bit_at(static_cast<const uint8_t *>(a->buffers[0]), a->offset)) return false;
        *v = value_at(s->field->format[0], a->buffers[1], a->offset);
        return true;
    }
    return false;
}

void release_inputs(pqb_series_export *in, size_t n) {        // the callee consumes its inputs (polars forgets them)
    for (size_t i = 0; i < n; ++i) {
        pqb_series_export *s = &in[i];
        if (s->arrays)
            for (size_t k = 0; k < s->len; ++k)
                if (s->arrays[k] && s->arrays[k]->release) s->arrays[k]->release(s->arrays[k]);
        if (s->release) s->release(s);
    }
}

// ---- kwargs: the pickle of a flat {str: int | float | bool | None} dict (what polars sends) ---------
struct Kw { std::string key; double val; bool none; };

bool parse_pickle(const uint8_t *p, size_t n, std::vector<Kw> *out) {
    struct Item { int kind; std::string s; double v; };      // kind: 0 str, 1 number, 2 none, 3 dict, 4 mark
    std::vector<Item> st;
    auto need = [&](size_t i, size_t k) { return i + k <= n; };
    size_t i = 0;
    auto flush_pairs = [&](size_t from) {
        for (size_t j = from; j + 1 < st.size(); j += 2)
            if (st[j].kind == 0) out->push_back({st[j].s, st[j + 1].v, st[j + 1].kind == 2});
        st.resize(from);
    };
    while (i < n) {
        const uint8_t op = p[i++];
        switch (op) {
            case 0x80: if (!need(i, 1)) return false; i += 1; break;                    // PROTO
            case 0x95: if (!need(i, 8)) return false; i += 8; break;                    // FRAME
            case 0x94: break;                                                           // MEMOIZE
            case 'q': if (!need(i, 1)) return false; i += 1; break;                     // BINPUT
            case 'r': if (!need(i, 4)) return false; i += 4; break;                     // LONG_BINPUT
            case '}': st.push_back({3, "", 0}); break;                                  // EMPTY_DICT
            case '(': st.push_back({4, "", 0}); break;                                  // MARK
            case 0x8c: {                                                                // SHORT_BINUNICODE
                if (!need(i, 1)) return false;
                const size_t len = p[i++];
                if (!need(i, len)) return false;
                st.push_back({0, std::string((const char *)p + i, len), 0});
                i += len;
                break;
            }
            case 'X': {                                                                 // BINUNICODE
                if (!need(i, 4)) return false;
                uint32_t len; memcpy(&len, p + i, 4); i += 4;
                if (!need(i, len)) return false;
                st.push_back({0, std::string((const char *)p + i, len), 0});
                i += len;
                break;
            }
            case 'K': if (!need(i, 1)) return false; st.push_back({1, "", (double)p[i]}); i += 1; break;           // BININT1
            case 'M': { if (!need(i, 2)) return false; uint16_t v; memcpy(&v, p + i, 2); st.push_back({1, "", (double)v}); i += 2; break; }
            case 'J': { if (!need(i, 4)) return false; int32_t v; memcpy(&v, p + i, 4); st.push_back({1, "", (double)v}); i += 4; break; }
            case 0x8a: {                                                                // LONG1
                if (!need(i, 1)) return false;
                const size_t len = p[i++];
                if (!need(i, len) || len > 8) return false;
                int64_t v = 0;
                for (size_t b = 0; b < len; ++b) v |= (int64_t)p[i + b] << (8 * b);
                if (len && len < 8 && (p[i + len - 1] & 0x80)) v -= (int64_t)1 << (8 * len);
                st.push_back({1, "", (double)v});
                i += len;
                break;
            }
            case 'G': {                                                                 // BINFLOAT (big endian)
                if (!need(i, 8)) return false;
                uint64_t b = 0;
                for (int k = 0; k < 8; ++k) b = (b << 8) | p[i + k];
                double v; memcpy(&v, &b, 8);
                st.push_back({1, "", v});
                i += 8;
                break;
            }
            case 0x88: st.push_back({1, "", 1.0}); break;                               // NEWTRUE
            case 0x89: st.push_back({1, "", 0.0}); break;                               // NEWFALSE
            case 'N': st.push_back({2, "", 0}); break;                                  // NONE
            case 's': {                                                                 // SETITEM
                if (st.size() < 3) return false;
                flush_pairs(st.size() - 2);
                break;
            }
            case 'u': {                                                                 // SETITEMS
                size_t m = st.size();
                while (m > 0 && st[m - 1].kind != 4) --m;
                if (m == 0) return false;
                flush_pairs(m);
                st.pop_back();                                                          // the mark
                break;
            }
            case '.': return true;                                                      // STOP
            default: return false;
        }
    }
    return true;
}

// ---- outputs --------------------------------------------------------------------------------------
struct PrimPriv { double *vals; uint8_t *valid; const void *bufs[2]; };

void release_prim(ArrowArray *a) {
    if (!a || !a->release) return;
    PrimPriv *p = static_cast<PrimPriv *>(a->private_data);
    free(p->vals);
    free(p->valid);
    delete p;
    a->release = nullptr;
}

struct StructPriv { const void *bufs[1]; std::vector<ArrowArray *> kids; };

void release_struct(ArrowArray *a) {
    if (!a || !a->release) return;
    StructPriv *p = static_cast<StructPriv *>(a->private_data);
    for (ArrowArray *k : p->kids) {
        if (k->release) k->release(k);
        delete k;
    }
    delete p;
    a->release = nullptr;
}

struct OutCol { double *vals = nullptr; uint8_t *valid = nullptr; int64_t n = 0; };

bool alloc_out(int64_t n, OutCol *o) {
    o->n = n;
    const size_t vb = (size_t)((std::max<int64_t>(n, 1) * 8 + 63) / 64 * 64);
    const size_t mb = (size_t)(((std::max<int64_t>(n, 1) + 7) / 8 + 63) / 64 * 64);
    void *a = nullptr, *b = nullptr;
    if (posix_memalign(&a, 64, vb) || posix_memalign(&b, 64, mb)) { free(a); set_err("out of memory"); return false; }
    memset(a, 0, vb);
    memset(b, 0, mb);
    o->vals = static_cast<double *>(a);
    o->valid = static_cast<uint8_t *>(b);
    return true;
}

void export_prim(OutCol &o, ArrowArray *a) {
    int64_t valid = 0;
    for (int64_t i = 0; i < (o.n + 7) / 8; ++i) valid += __builtin_popcount(o.valid[i]);
    PrimPriv *p = new PrimPriv{o.vals, o.valid, {o.valid, o.vals}};
    memset(a, 0, sizeof(*a));
    a->length = o.n;
    a->null_count = o.n - valid;
    a->n_buffers = 2;
    a->buffers = p->bufs;
    a->release = release_prim;
    a->private_data = p;
    o.vals = nullptr;
    o.valid = nullptr;
}

struct SchemaPriv { std::string name; std::vector<ArrowSchema *> kids; };

void release_schema(ArrowSchema *s) {
    if (!s || !s->release) return;
    SchemaPriv *p = static_cast<SchemaPriv *>(s->private_data);
    for (ArrowSchema *k : p->kids) {
        if (k->release) k->release(k);
        delete k;
    }
    delete p;
    s->release = nullptr;
}

void export_schema(ArrowSchema *s, const char *name, int n_fields, const char *const *field_names, const char *prim = "g") {
    SchemaPriv *p = new SchemaPriv{name, {}};
    memset(s, 0, sizeof(*s));
    s->format = n_fields ? "+s" : prim;
    s->name = p->name.c_str();
    s->flags = 2;                                             // ARROW_FLAG_NULLABLE
    for (int i = 0; i < n_fields; ++i) {
        ArrowSchema *k = new ArrowSchema;
        export_schema(k, field_names[i], 0, nullptr);
        p->kids.push_back(k);
    }
    s->n_children = n_fields;
    s->children = n_fields ? p->kids.data() : nullptr;
    s->release = release_schema;
    s->private_data = p;
}

struct SeriesPriv { ArrowSchema *field; ArrowArray **arrays; };

void release_series(pqb_series_export *e) {                   // polars-ffi protocol: the importer took the arrays
    if (!e || !e->release) return;
    SeriesPriv *p = static_cast<SeriesPriv *>(e->private_data);
    if (p->field->release) p->field->release(p->field);
    delete p->field;
    delete p->arrays[0];                                      // the struct's storage, not the array
    delete[] p->arrays;
    delete p;
    e->release = nullptr;
    e->private_data = nullptr;
}

// ---- the function table ---------------------------------------------------------------------------
struct Spec {
    const char *name;
    int n_cols;
    int n_params;
    const char *pname[6];
    double pdef[6];
    int n_out;                        // 1 = Float64, >1 = struct
    const char *struct_name;
    const char *out_names[3];
};

enum Fn { F_SMA, F_EMA, F_TEMA, F_TRIMA, F_MA, F_BBANDS, F_MIDPOINT, F_MIDPRICE, F_RSI, F_MACD, F_WILLR, F_MOM,
          F_ROC, F_ROCP, F_ROCR, F_ROCR100, F_CMO, F_MFI, F_CCI, F_TRANGE, F_ATR, F_NATR, F_OBV, F_AD, F_ADOSC,
          F_STOCH, F_KDJ, F_STOCHF, F_STOCHRSI, F_MACDEXT, F_ADX, F_ADXR, F_DX, F_PLUS_DI, F_MINUS_DI, F_PLUS_DM, F_MINUS_DM, F_TRIX, F_ULTOSC, F_AROON, F_COUNT };

const Spec SPECS[F_COUNT] = {
    {"sma", 1, 1, {"timeperiod"}, {30}, 1, nullptr, {}},
    {"ema", 1, 1, {"timeperiod"}, {30}, 1, nullptr, {}},
    {"tema", 1, 1, {"timeperiod"}, {30}, 1, nullptr, {}},
    {"trima", 1, 1, {"timeperiod"}, {30}, 1, nullptr, {}},
    {"ma", 1, 2, {"timeperiod", "matype"}, {30, 0}, 1, nullptr, {}},
    {"bbands", 1, 3, {"timeperiod", "nbdevup", "nbdevdn"}, {20, 2.0, 2.0}, 3, "bbands", {"bb_upper", "bb_middle", "bb_lower"}},
    {"midpoint", 1, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"midprice", 2, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"rsi", 1, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"macd", 1, 3, {"fastperiod", "slowperiod", "signalperiod"}, {12, 26, 9}, 3, "macd_res", {"macd", "macd_signal", "macd_hist"}},
    {"willr", 3, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"mom", 1, 1, {"timeperiod"}, {10}, 1, nullptr, {}},
    {"roc", 1, 1, {"timeperiod"}, {10}, 1, nullptr, {}},
    {"rocp", 1, 1, {"timeperiod"}, {10}, 1, nullptr, {}},
    {"rocr", 1, 1, {"timeperiod"}, {10}, 1, nullptr, {}},
    {"rocr100", 1, 1, {"timeperiod"}, {10}, 1, nullptr, {}},
    {"cmo", 1, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"mfi", 4, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"cci", 3, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"trange", 3, 0, {}, {}, 1, nullptr, {}},
    {"atr", 3, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"natr", 3, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"obv", 2, 0, {}, {}, 1, nullptr, {}},
    {"ad", 4, 0, {}, {}, 1, nullptr, {}},
    {"adosc", 4, 2, {"fastperiod", "slowperiod"}, {3, 10}, 1, nullptr, {}},
    {"stoch", 3, 5, {"fastk_period", "slowk_period", "slowk_matype", "slowd_period", "slowd_matype"}, {5, 3, 0, 3, 0}, 2, "stoch", {"slowk", "slowd"}},
    {"kdj", 3, 3, {"fastk_period", "k_period", "d_period"}, {9, 3, 3}, 3, "kdj", {"k", "d", "j"}},
    {"stochf", 3, 3, {"fastk_period", "fastd_period", "fastd_matype"}, {5, 3, 0}, 2, "stochf", {"fastk", "fastd"}},
    {"stochrsi", 1, 4, {"timeperiod", "fastk_period", "fastd_period", "fastd_matype"}, {14, 5, 3, 0}, 2, "stochrsi", {"fastk_rsi", "fastd_rsi"}},
    {"macdext", 1, 6, {"fastperiod", "fastmatype", "slowperiod", "slowmatype", "signalperiod", "signalmatype"}, {12, 0, 26, 0, 9, 0}, 3, "macdext", {"macd_dif", "macd_dea", "macd_hist"}},
    {"adx", 3, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"adxr", 3, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"dx", 3, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"plus_di", 3, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"minus_di", 3, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"plus_dm", 2, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"minus_dm", 2, 1, {"timeperiod"}, {14}, 1, nullptr, {}},
    {"trix", 1, 1, {"timeperiod"}, {30}, 1, nullptr, {}},
    {"ultosc", 3, 3, {"timeperiod1", "timeperiod2", "timeperiod3"}, {7, 14, 28}, 1, nullptr, {}},
    {"aroon", 2, 1, {"timeperiod"}, {14}, 2, "aroon", {"aroon_up", "aroon_down"}},
};

int run_fn(int fn, pqb_engine *e, const pqb_col *c, const double *pv, pqb_out_col *o) {
    const int32_t p0 = (int32_t)pv[0], p1 = (int32_t)pv[1], p2 = (int32_t)pv[2], p3 = (int32_t)pv[3], p4 = (int32_t)pv[4], p5 = (int32_t)pv[5];
    switch (fn) {
        case F_SMA: return pqb_sma(e, &c[0], p0, &o[0]);
        case F_EMA: return pqb_ema(e, &c[0], p0, &o[0]);
        case F_TEMA: return pqb_tema(e, &c[0], p0, &o[0]);
        case F_TRIMA: return pqb_trima(e, &c[0], p0, &o[0]);
        case F_MA: return pqb_ma(e, &c[0], p0, p1, &o[0]);
        case F_BBANDS: return pqb_bbands(e, &c[0], p0, pv[1], pv[2], &o[0], &o[1], &o[2]);
        case F_MIDPOINT: return pqb_midpoint(e, &c[0], p0, &o[0]);
        case F_MIDPRICE: return pqb_midprice(e, &c[0], &c[1], p0, &o[0]);
        case F_RSI: return pqb_rsi(e, &c[0], p0, &o[0]);
        case F_MACD: return pqb_macd(e, &c[0], p0, p1, p2, &o[0], &o[1], &o[2]);
        case F_WILLR: return pqb_willr(e, &c[0], &c[1], &c[2], p0, &o[0]);
        case F_MOM: return pqb_mom(e, &c[0], p0, &o[0]);
        case F_ROC: return pqb_roc(e, &c[0], p0, 0, &o[0]);
        case F_ROCP: return pqb_roc(e, &c[0], p0, 1, &o[0]);
        case F_ROCR: return pqb_roc(e, &c[0], p0, 2, &o[0]);
        case F_ROCR100: return pqb_roc(e, &c[0], p0, 3, &o[0]);
        case F_CMO: return pqb_cmo(e, &c[0], p0, &o[0]);
        case F_MFI: return pqb_mfi(e, &c[0], &c[1], &c[2], &c[3], p0, &o[0]);
        case F_CCI: return pqb_cci(e, &c[0], &c[1], &c[2], p0, &o[0]);
        case F_TRANGE: return pqb_trange(e, &c[0], &c[1], &c[2], &o[0]);
        case F_ATR: return pqb_atr(e, &c[0], &c[1], &c[2], p0, &o[0]);
        case F_NATR: return pqb_natr(e, &c[0], &c[1], &c[2], p0, &o[0]);
        case F_OBV: return pqb_obv(e, &c[0], &c[1], &o[0]);
        case F_AD: return pqb_ad(e, &c[0], &c[1], &c[2], &c[3], &o[0]);
        case F_ADOSC: return pqb_adosc(e, &c[0], &c[1], &c[2], &c[3], p0, p1, &o[0]);
        case F_STOCH: return pqb_stoch_ma(e, &c[0], &c[1], &c[2], p0, p1, p2, p3, p4, &o[0], &o[1]);
        case F_STOCHF: return pqb_stochf(e, &c[0], &c[1], &c[2], p0, p1, p2, &o[0], &o[1]);
        case F_STOCHRSI: return pqb_stochrsi(e, &c[0], p0, p1, p2, p3, &o[0], &o[1]);
        case F_MACDEXT: return pqb_macdext(e, &c[0], p0, p1, p2, p3, p4, p5, &o[0], &o[1], &o[2]);
        case F_KDJ: return pqb_kdj(e, &c[0], &c[1], &c[2], p0, p1, p2, &o[0], &o[1], &o[2]);
        // the directional-movement family: one fused call, one requested output (momentum.rs:11-61, 226-237, 344-436)
        case F_ADX: return pqb_dm(e, &c[0], &c[1], &c[2], p0, nullptr, nullptr, nullptr, nullptr, &o[0], nullptr);
        case F_ADXR: return pqb_dm(e, &c[0], &c[1], &c[2], p0, nullptr, nullptr, nullptr, nullptr, nullptr, &o[0]);
        case F_DX: return pqb_dm(e, &c[0], &c[1], &c[2], p0, nullptr, nullptr, &o[0], nullptr, nullptr, nullptr);
        case F_PLUS_DI: return pqb_dm(e, &c[0], &c[1], &c[2], p0, nullptr, nullptr, &o[0], nullptr, nullptr, nullptr);   // :409 returns calc_dm().0 = DX
        case F_MINUS_DI: return pqb_dm(e, &c[0], &c[1], &c[2], p0, nullptr, nullptr, nullptr, &o[0], nullptr, nullptr);
        case F_PLUS_DM: return pqb_dm(e, &c[0], &c[1], nullptr, p0, &o[0], nullptr, nullptr, nullptr, nullptr, nullptr);
        case F_MINUS_DM: return pqb_dm(e, &c[0], &c[1], nullptr, p0, nullptr, &o[0], nullptr, nullptr, nullptr, nullptr);
        case F_TRIX: return pqb_trix(e, &c[0], p0, &o[0]);
        case F_ULTOSC: return pqb_ultosc(e, &c[0], &c[1], &c[2], p0, p1, p2, &o[0]);
        case F_AROON: return pqb_aroon(e, &c[0], &c[1], p0, &o[0], &o[1]);
    }
    return PQB_ERR_INVALID;
}

bool call_impl(int fn, pqb_series_export *in, size_t n_in, const uint8_t *kw, size_t kwl, pqb_series_export *ret) {
    const Spec &S = SPECS[fn];
    if (!in || n_in < (size_t)S.n_cols) { set_err("%s: expected %d input columns, got %zu", S.name, S.n_cols, n_in); return false; }
    if (!ret) { set_err("%s: NULL return_value", S.name); return false; }
    // parameters: defaults <- trailing literal inputs (the Python shims) <- pickled kwargs (the Rust structs)
    double pv[6] = {S.pdef[0], S.pdef[1], S.pdef[2], S.pdef[3], S.pdef[4], S.pdef[5]};
    for (int k = 0; k < S.n_params; ++k)
        if ((size_t)(S.n_cols + k) < n_in) {
            double v;
            if (literal_value(&in[S.n_cols + k], &v)) pv[k] = v;
        }
    if (kw && kwl) {
        std::vector<Kw> kws;
        if (!parse_pickle(kw, kwl, &kws)) { set_err("%s: cannot decode the kwargs pickle (%zu bytes)", S.name, kwl); return false; }
        for (const Kw &k : kws) {
            bool known = false;
            for (int j = 0; j < S.n_params; ++j)
                if (k.key == S.pname[j]) { known = true; if (!k.none) pv[j] = k.val; }
            if (!known) { set_err("%s: unknown kwarg '%s'", S.name, k.key.c_str()); return false; }
        }
    }
    for (int k = 0; k < S.n_params; ++k) {
        const bool is_float = (fn == F_BBANDS && k > 0);
        if (!is_float && (!(pv[k] >= 0) || pv[k] > 2147483647.0 || pv[k] != std::floor(pv[k]))) {
            set_err("%s: %s must be a non-negative integer (usize in the reference), got %g", S.name, S.pname[k], pv[k]);
            return false;
        }
    }
    Column cols[4];
    for (int i = 0; i < S.n_cols; ++i)
        if (!import_column(&in[i], i, &cols[i])) return false;
    const int64_t n = cols[0].col.len;
    for (int i = 1; i < S.n_cols; ++i)
        if (cols[i].col.len != n) { set_err("%s: input columns differ in length (%lld vs %lld)", S.name, (long long)n, (long long)cols[i].col.len); return false; }
    OutCol outs[3];
    for (int k = 0; k < S.n_out; ++k)
        if (!alloc_out(n, &outs[k])) { for (int j = 0; j < k; ++j) { free(outs[j].vals); free(outs[j].valid); } return false; }
    bool ok = true;
    if (n > 0) {
        pqb_engine *e = engine();
        if (!e) ok = false;
        if (ok) {
            pqb_col c[4];
            pqb_out_col o[3];
            for (int i = 0; i < S.n_cols; ++i) c[i] = cols[i].col;
            for (int k = 0; k < S.n_out; ++k) o[k] = pqb_out_col{outs[k].vals, outs[k].valid};
            const int rc = run_fn(fn, e, c, pv, o);
            if (rc != PQB_OK) { set_err("%s: %s", S.name, pqb_last_error()); ok = false; }
        }
    }
    if (!ok) {
        for (int k = 0; k < S.n_out; ++k) { free(outs[k].vals); free(outs[k].valid); }
        return false;
    }
    // export: Float64 named after the first input (FieldsMapper::with_dtype), or the reference's struct
    ArrowArray *arr = new ArrowArray;
    ArrowSchema *field = new ArrowSchema;
    if (S.n_out == 1) {
        export_prim(outs[0], arr);
        export_schema(field, cols[0].name.c_str(), 0, nullptr);
    } else {
        StructPriv *sp = new StructPriv{{nullptr}, {}};
        for (int k = 0; k < S.n_out; ++k) {
            ArrowArray *kid = new ArrowArray;
            export_prim(outs[k], kid);
            sp->kids.push_back(kid);
        }
        memset(arr, 0, sizeof(*arr));
        arr->length = n;
        arr->n_buffers = 1;
        arr->buffers = sp->bufs;
        arr->n_children = S.n_out;
        arr->children = sp->kids.data();
        arr->release = release_struct;
        arr->private_data = sp;
        export_schema(field, S.struct_name, S.n_out, S.out_names);
    }
    SeriesPriv *pr = new SeriesPriv{field, new ArrowArray *[1]{arr}};
    ret->field = field;
    ret->arrays = pr->arrays;
    ret->len = 1;
    ret->release = release_series;
    ret->private_data = pr;
    return true;
}

// ---- candle functions: cdl* (Int32) and the price transforms / bop -----------------------------------
struct I32Priv { int32_t *vals; const void *bufs[2]; };

void release_i32(ArrowArray *a) {
    if (!a || !a->release) return;
    I32Priv *p = static_cast<I32Priv *>(a->private_data);
    free(p->vals);
    delete p;
    a->release = nullptr;
}

void finish_series(pqb_series_export *ret, ArrowArray *arr, ArrowSchema *field) {
    SeriesPriv *pr = new SeriesPriv{field, new ArrowArray *[1]{arr}};
    ret->field = field;
    ret->arrays = pr->arrays;
    ret->len = 1;
    ret->release = release_series;
    ret->private_data = pr;
}

// pattern.rs:9 `#[polars_expr(output_type=Int32)] pub fn cdl2crows(inputs)`: open, high, low, close (+ an optional
// Float64 `penetration` literal, `inputs.get(4)`); the Series is named after the function (`from_slice("cdl2crows"...)`)
bool cdl_impl(int pattern, pqb_series_export *in, size_t n_in, pqb_series_export *ret) {
    const char *name = pqb_pattern_name(pattern);
    if (!in || n_in < 4) { set_err("%s: expected 4 input columns (open, high, low, close), got %zu", name, n_in); return false; }
    if (!ret) { set_err("%s: NULL return_value", name); return false; }
    double pen = 0.3;
    if (n_in > 4) literal_value(&in[4], &pen);
    Column cols[4];
    for (int i = 0; i < 4; ++i)
        if (!import_column(&in[i], i, &cols[i])) return false;
    const int64_t n = cols[0].col.len;
    for (int i = 1; i < 4; ++i)
        if (cols[i].col.len != n) { set_err("%s: input columns differ in length", name); return false; }
    void *buf = nullptr;
    if (posix_memalign(&buf, 64, (size_t)((std::max<int64_t>(n, 1) * 4 + 63) / 64 * 64))) { set_err("out of memory"); return false; }
    memset(buf, 0, (size_t)std::max<int64_t>(n, 1) * 4);
    if (n > 0) {
        pqb_engine *e = engine();
        int rc = e ? pqb_cdl(e, pattern, &cols[0].col, &cols[1].col, &cols[2].col, &cols[3].col, pen, static_cast<int32_t *>(buf))
                   : PQB_ERR_NO_DEVICE;
        if (rc != PQB_OK) {
            if (e) set_err("%s: %s", name, pqb_last_error());
            free(buf);
            return false;
        }
    } else {
        // the reference refuses nulls even in an empty column? cont_slice() of an empty array is fine: nothing to do
    }
    I32Priv *p = new I32Priv{static_cast<int32_t *>(buf), {nullptr, buf}};
    ArrowArray *arr = new ArrowArray;
    memset(arr, 0, sizeof(*arr));
    arr->length = n;
    arr->n_buffers = 2;
    arr->buffers = p->bufs;
    arr->release = release_i32;
    arr->private_data = p;
    ArrowSchema *field = new ArrowSchema;
    export_schema(field, name, 0, nullptr, "i");
    finish_series(ret, arr, field);
    return true;
}

// price.rs:10-91 avgprice(open, high, low, close) medprice(high, low) typprice / wclprice(high, low, close);
// momentum.rs:113 bop(open, high, low, close)
bool price_impl(int which, pqb_series_export *in, size_t n_in, pqb_series_export *ret) {
    static const char *const names[5] = {"avgprice", "medprice", "typprice", "wclprice", "bop"};
    static const int n_cols[5] = {4, 2, 3, 3, 4};
    static const int slot[5][4] = {{0, 1, 2, 3}, {1, 2, -1, -1}, {1, 2, 3, -1}, {1, 2, 3, -1}, {0, 1, 2, 3}};   // -> o, h, l, c
    const char *name = names[which];
    if (!in || n_in < (size_t)n_cols[which]) { set_err("%s: expected %d input columns, got %zu", name, n_cols[which], n_in); return false; }
    if (!ret) { set_err("%s: NULL return_value", name); return false; }
    Column cols[4];
    const pqb_col *ohlc[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t n = -1;
    for (int i = 0; i < n_cols[which]; ++i) {
        if (!import_column(&in[i], i, &cols[i])) return false;
        if (n >= 0 && cols[i].col.len != n) { set_err("%s: input columns differ in length", name); return false; }
        n = cols[i].col.len;
        ohlc[slot[which][i]] = &cols[i].col;
    }
    OutCol out;
    if (!alloc_out(n, &out)) return false;
    if (n > 0) {
        pqb_engine *e = engine();
        pqb_out_col o{out.vals, out.valid};
        int rc = e ? pqb_price(e, which, ohlc[0], ohlc[1], ohlc[2], ohlc[3], &o) : PQB_ERR_NO_DEVICE;
        if (rc != PQB_OK) {
            if (e) set_err("%s: %s", name, pqb_last_error());
            free(out.vals);
            free(out.valid);
            return false;
        }
    }
    ArrowArray *arr = new ArrowArray;
    export_prim(out, arr);
    ArrowSchema *field = new ArrowSchema;
    export_schema(field, name, 0, nullptr);
    finish_series(ret, arr, field);
    return true;
}

template <class F>
void guarded(const char *name, pqb_series_export *in, size_t n_in, F &&f) {
    g_perr.clear();
    try {
        f();
    } catch (const std::exception &ex) {
        set_err("%s: %s", name, ex.what());
    } catch (...) {
        set_err("%s: unknown C++ exception", name);
    }
    if (in) release_inputs(in, n_in);
}

void typed_field(ArrowSchema *fields, size_t n, ArrowSchema *out, const char *fmt) {       // FieldsMapper::with_dtype
    g_perr.clear();
    if (!out) return;
    try {
        export_schema(out, (fields && n && fields[0].name) ? fields[0].name : "", 0, nullptr, fmt);
    } catch (...) {
        set_err("cannot build the output field");
    }
}

void plugin_call(int fn, pqb_series_export *in, size_t n_in, const uint8_t *kw, size_t kwl, pqb_series_export *ret) {
    g_perr.clear();
    bool ok = false;
    try {
        ok = call_impl(fn, in, n_in, kw, kwl, ret);
    } catch (const std::exception &ex) {                     // never throw across the ABI
        set_err("%s: %s", SPECS[fn].name, ex.what());
    } catch (...) {
        set_err("%s: unknown C++ exception", SPECS[fn].name);
    }
    (void)ok;
    if (in) release_inputs(in, n_in);
}

void plugin_field(int fn, ArrowSchema *fields, size_t n, ArrowSchema *out) {
    g_perr.clear();
    if (!out) return;
    const Spec &S = SPECS[fn];
    try {
        if (S.n_out == 1) export_schema(out, (fields && n && fields[0].name) ? fields[0].name : "", 0, nullptr);
        else export_schema(out, S.struct_name, S.n_out, S.out_names);
    } catch (...) {
        set_err("%s: cannot build the output field", S.name);
    }
}

}  // namespace

extern "C" uint32_t _polars_plugin_get_version(void) { return (0u << 16) | 1u; }
extern "C" const char *_polars_plugin_get_last_error_message(void) { return g_perr.c_str(); }

#define PQB_DEFINE_PLUGIN(name, id)                                                                            \
    extern "C" void _polars_plugin_##name(pqb_series_export *in, size_t n, const uint8_t *kw, size_t kwl,      \
                                          pqb_series_export *ret, void *) { plugin_call(id, in, n, kw, kwl, ret); } \
    extern "C" void _polars_plugin_field_##name(ArrowSchema *f, size_t n, ArrowSchema *out) { plugin_field(id, f, n, out); }

PQB_DEFINE_PLUGIN(sma, F_SMA)
PQB_DEFINE_PLUGIN(ema, F_EMA)
PQB_DEFINE_PLUGIN(tema, F_TEMA)
PQB_DEFINE_PLUGIN(trima, F_TRIMA)
PQB_DEFINE_PLUGIN(ma, F_MA)
PQB_DEFINE_PLUGIN(bbands, F_BBANDS)
PQB_DEFINE_PLUGIN(midpoint, F_MIDPOINT)
PQB_DEFINE_PLUGIN(midprice, F_MIDPRICE)
PQB_DEFINE_PLUGIN(rsi, F_RSI)
PQB_DEFINE_PLUGIN(macd, F_MACD)
PQB_DEFINE_PLUGIN(willr, F_WILLR)
PQB_DEFINE_PLUGIN(mom, F_MOM)
PQB_DEFINE_PLUGIN(roc, F_ROC)
PQB_DEFINE_PLUGIN(rocp, F_ROCP)
PQB_DEFINE_PLUGIN(rocr, F_ROCR)
PQB_DEFINE_PLUGIN(rocr100, F_ROCR100)
PQB_DEFINE_PLUGIN(cmo, F_CMO)
PQB_DEFINE_PLUGIN(mfi, F_MFI)
PQB_DEFINE_PLUGIN(cci, F_CCI)
PQB_DEFINE_PLUGIN(trange, F_TRANGE)
PQB_DEFINE_PLUGIN(atr, F_ATR)
PQB_DEFINE_PLUGIN(natr, F_NATR)
PQB_DEFINE_PLUGIN(obv, F_OBV)
PQB_DEFINE_PLUGIN(ad, F_AD)
PQB_DEFINE_PLUGIN(adosc, F_ADOSC)
PQB_DEFINE_PLUGIN(stoch, F_STOCH)
PQB_DEFINE_PLUGIN(kdj, F_KDJ)
PQB_DEFINE_PLUGIN(stochf, F_STOCHF)
PQB_DEFINE_PLUGIN(stochrsi, F_STOCHRSI)
PQB_DEFINE_PLUGIN(macdext, F_MACDEXT)
PQB_DEFINE_PLUGIN(adx, F_ADX)
PQB_DEFINE_PLUGIN(adxr, F_ADXR)
PQB_DEFINE_PLUGIN(dx, F_DX)
PQB_DEFINE_PLUGIN(plus_di, F_PLUS_DI)
PQB_DEFINE_PLUGIN(minus_di, F_MINUS_DI)
PQB_DEFINE_PLUGIN(plus_dm, F_PLUS_DM)
PQB_DEFINE_PLUGIN(minus_dm, F_MINUS_DM)
PQB_DEFINE_PLUGIN(trix, F_TRIX)
PQB_DEFINE_PLUGIN(ultosc, F_ULTOSC)
PQB_DEFINE_PLUGIN(aroon, F_AROON)

// ---- candle symbols: the 61 cdl* functions (pattern.rs) in the reference's order, price.rs, bop ---------
#define PQB_DEFINE_CDL(name, id)                                                                               \
    extern "C" void _polars_plugin_##name(pqb_series_export *in, size_t n, const uint8_t *, size_t,            \
                                          pqb_series_export *ret, void *) {                                    \
        guarded(#name, in, n, [&] { cdl_impl(id, in, n, ret); });                                              \
    }                                                                                                          \
    extern "C" void _polars_plugin_field_##name(ArrowSchema *f, size_t n, ArrowSchema *out) { typed_field(f, n, out, "i"); }
#define PQB_DEFINE_PRICE(name, id)                                                                             \
    extern "C" void _polars_plugin_##name(pqb_series_export *in, size_t n, const uint8_t *, size_t,            \
                                          pqb_series_export *ret, void *) {                                    \
        guarded(#name, in, n, [&] { price_impl(id, in, n, ret); });                                            \
    }                                                                                                          \
    extern "C" void _polars_plugin_field_##name(ArrowSchema *f, size_t n, ArrowSchema *out) { typed_field(f, n, out, "g"); }

#define X(n) PQB_DEFINE_CDL(n, pqb_pattern_index(#n))
PQB_CDL_LIST(X)
#undef X
PQB_DEFINE_PRICE(avgprice, PQB_PRICE_AVG)
PQB_DEFINE_PRICE(medprice, PQB_PRICE_MED)
PQB_DEFINE_PRICE(typprice, PQB_PRICE_TYP)
PQB_DEFINE_PRICE(wclprice, PQB_PRICE_WCL)
PQB_DEFINE_PRICE(bop, PQB_PRICE_BOP)
