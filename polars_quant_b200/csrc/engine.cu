// engine.cu -- host side of libpqb200.so: the C ABI of include/pqb200.h over the fused
// sm_100a suite kernel (suite_kernel.cuh).  Plain CUDA runtime; no torch, no CPU fallback:
// every compute entry point fails loudly when there is no device.
#include "../../include/pqb200.h"
#include "suite_kernel.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace pqb;

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            int code__ = (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver)         \
                             ? PQB_ERR_NO_DEVICE : PQB_ERR_CUDA;                                  \
            return fail(code__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                         \
    } while (0)

extern "C" const char *pqb_last_error(void) { return g_err.c_str(); }
extern "C" int pqb_abi_version(void) { return PQB_ABI_VERSION; }

extern "C" int pqb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int good = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp pr;
        if (cudaGetDeviceProperties(&pr, d) == cudaSuccess && pr.major == 10) ++good;
    }
    return good;
}

extern "C" void pqb_suite_params_default(pqb_suite_params *p) {
    if (!p) return;
    p->indicators = PQB_IND_ALL;
    p->sma_period = 30; p->ema_period = 30; p->tema_period = 30; p->trima_period = 30;
    p->bbands_period = 20; p->bbands_nbdevup = 2.0; p->bbands_nbdevdn = 2.0;
    p->macd_fast = 12; p->macd_slow = 26; p->macd_signal = 9;
    p->rsi_period = 14; p->atr_period = 14; p->natr_period = 14;
    p->kdj_fastk = 9; p->kdj_slowk = 3; p->kdj_slowd = 3;
    p->willr_period = 14; p->midprice_period = 14;
}

// ---------------------------------------------------------------------------------------
// engine / panel objects
// ---------------------------------------------------------------------------------------
struct pqb_engine {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t h2d = nullptr, d2h = nullptr;
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
    int ctas_per_sm32 = 0, ctas_per_sm128 = 0;
    std::mutex mu;                       // guards the single-column scratch panel
    pqb_panel *scratch = nullptr;
    int64_t scratch_bars = 0;
};

struct pqb_panel {
    pqb_engine *e = nullptr;
    int64_t n_symbols = 0, n_bars = 0, pitch = 0, words_per_row = 0;
    uint32_t fields_mask = 0, outputs_mask = 0;
    double *d_in[PQB_N_FIELDS] = {};
    double *d_out[PQB_N_OUTPUTS] = {};
    uint32_t *d_bits[PQB_N_OUTPUTS] = {};
    int *d_start = nullptr;
    bool starts_nonzero = false;
    // pinned staging
    double *h_in[PQB_N_FIELDS] = {};
    double *h_out[PQB_N_OUTPUTS] = {};
    uint32_t *h_bits[PQB_N_OUTPUTS] = {};
    std::vector<int32_t> h_start;
    bool staging = false;
    cudaEvent_t ev[4] = {};
};

static int set_dev(const pqb_engine *e) {
    CU(cudaSetDevice(e->device));
    return PQB_OK;
}

static constexpr int kWarpsPerCta = PQB_CTA_THREADS / 32;

template <int HALO>
static int configure_kernel(int *ctas_per_sm) {
    const int smem = kWarpsPerCta * WarpSmem<HALO>::BYTES;
    CU(cudaFuncSetAttribute(suite_fused_kernel<HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(suite_fused_kernel<HALO>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int n = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, suite_fused_kernel<HALO>, PQB_CTA_THREADS, smem));
    *ctas_per_sm = std::max(n, 1);
    return PQB_OK;
}

extern "C" int pqb_engine_create(int device, pqb_engine **out) {
    if (!out) return fail(PQB_ERR_INVALID, "pqb_engine_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(PQB_ERR_NO_DEVICE, "no CUDA device (%s): the engine has no CPU fallback",
                    ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    }
    if (device < 0 || device >= n) return fail(PQB_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    cudaDeviceProp pr;
    CU(cudaGetDeviceProperties(&pr, device));
    if (pr.major != 10)
        return fail(PQB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    pr.major, pr.minor);
    pqb_engine *e = new pqb_engine();
    e->device = device;
    e->sm_count = pr.multiProcessorCount;
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&e->h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&e->d2h, cudaStreamNonBlocking));
    int rc = configure_kernel<32>(&e->ctas_per_sm32);
    if (rc) { delete e; return rc; }
    rc = configure_kernel<128>(&e->ctas_per_sm128);
    if (rc) { delete e; return rc; }
    *out = e;
    return PQB_OK;
}

extern "C" void pqb_engine_destroy(pqb_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->scratch) pqb_panel_destroy(e->scratch);
    if (e->flush_buf) cudaFree(e->flush_buf);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->h2d) cudaStreamDestroy(e->h2d);
    if (e->d2h) cudaStreamDestroy(e->d2h);
    delete e;
}

extern "C" int pqb_panel_create(pqb_engine *e, int64_t n_symbols, int64_t n_bars, uint32_t fields_mask,
                                uint32_t outputs_mask, int host_staging, pqb_panel **out) {
    if (!e || !out) return fail(PQB_ERR_INVALID, "pqb_panel_create: NULL argument");
    *out = nullptr;
    if (n_symbols <= 0 || n_bars <= 0 || n_symbols > (1ll << 30) || n_bars > (1ll << 30))
        return fail(PQB_ERR_INVALID, "pqb_panel_create: bad shape %lld x %lld", (long long)n_symbols, (long long)n_bars);
    if (fields_mask == 0 || fields_mask >= (1u << PQB_N_FIELDS) || outputs_mask >= (1u << PQB_N_OUTPUTS))
        return fail(PQB_ERR_INVALID, "pqb_panel_create: bad masks");
    int rc = set_dev(e);
    if (rc) return rc;
    pqb_panel *p = new pqb_panel();
    p->e = e;
    p->n_symbols = n_symbols;
    p->n_bars = n_bars;
    p->pitch = (n_bars + 15) / 16 * 16;
    p->words_per_row = (n_bars + 31) / 32;
    p->fields_mask = fields_mask;
    p->outputs_mask = outputs_mask;
    p->staging = host_staging != 0;
    const size_t plane = (size_t)n_symbols * p->pitch * sizeof(double);
    const size_t bplane = (size_t)n_symbols * p->words_per_row * sizeof(uint32_t);
    auto bail = [&](cudaError_t ce, const char *what) {
        fail(ce == cudaErrorMemoryAllocation ? PQB_ERR_ALLOC : PQB_ERR_CUDA, "pqb_panel_create: %s: %s", what,
             cudaGetErrorString(ce));
        pqb_panel_destroy(p);
        return g_err.empty() ? PQB_ERR_CUDA : (ce == cudaErrorMemoryAllocation ? PQB_ERR_ALLOC : PQB_ERR_CUDA);
    };
    cudaError_t ce;
    for (int f = 0; f < PQB_N_FIELDS; ++f) {
        if (!(fields_mask >> f & 1)) continue;
        // + one tile of slack: the last TMA tile of the last row never reads past the allocation
        if ((ce = cudaMalloc(&p->d_in[f], plane + TILE * sizeof(double))) != cudaSuccess) return bail(ce, "cudaMalloc(field)");
        if ((ce = cudaMemsetAsync(p->d_in[f], 0, plane + TILE * sizeof(double), e->stream)) != cudaSuccess) return bail(ce, "memset");
        if (p->staging) {
            if ((ce = cudaMallocHost(&p->h_in[f], plane)) != cudaSuccess) return bail(ce, "cudaMallocHost(field)");
            memset(p->h_in[f], 0, plane);
        }
    }
    for (int k = 0; k < PQB_N_OUTPUTS; ++k) {
        if (!(outputs_mask >> k & 1)) continue;
        if ((ce = cudaMalloc(&p->d_out[k], plane)) != cudaSuccess) return bail(ce, "cudaMalloc(output)");
        if ((ce = cudaMalloc(&p->d_bits[k], bplane)) != cudaSuccess) return bail(ce, "cudaMalloc(validity)");
        if (p->staging) {
            if ((ce = cudaMallocHost(&p->h_out[k], plane)) != cudaSuccess) return bail(ce, "cudaMallocHost(output)");
            if ((ce = cudaMallocHost(&p->h_bits[k], bplane)) != cudaSuccess) return bail(ce, "cudaMallocHost(validity)");
        }
    }
    if ((ce = cudaMalloc(&p->d_start, (size_t)n_symbols * sizeof(int))) != cudaSuccess) return bail(ce, "cudaMalloc(start)");
    if ((ce = cudaMemsetAsync(p->d_start, 0, (size_t)n_symbols * sizeof(int), e->stream)) != cudaSuccess) return bail(ce, "memset");
    p->h_start.assign((size_t)n_symbols, 0);
    for (auto &ev : p->ev)
        if ((ce = cudaEventCreate(&ev)) != cudaSuccess) return bail(ce, "cudaEventCreate");
    if ((ce = cudaStreamSynchronize(e->stream)) != cudaSuccess) return bail(ce, "sync");
    *out = p;
    return PQB_OK;
}

extern "C" void pqb_panel_destroy(pqb_panel *p) {
    if (!p) return;
    if (p->e) cudaSetDevice(p->e->device);
    for (auto &q : p->d_in) if (q) cudaFree(q);
    for (auto &q : p->d_out) if (q) cudaFree(q);
    for (auto &q : p->d_bits) if (q) cudaFree(q);
    if (p->d_start) cudaFree(p->d_start);
    for (auto &q : p->h_in) if (q) cudaFreeHost(q);
    for (auto &q : p->h_out) if (q) cudaFreeHost(q);
    for (auto &q : p->h_bits) if (q) cudaFreeHost(q);
    for (auto &ev : p->ev) if (ev) cudaEventDestroy(ev);
    delete p;
}

extern "C" int64_t pqb_panel_pitch(const pqb_panel *p) { return p ? p->pitch : 0; }
extern "C" int64_t pqb_panel_validity_pitch(const pqb_panel *p) { return p ? p->words_per_row * 4 : 0; }
extern "C" double *pqb_panel_host_field(pqb_panel *p, int f) {
    return (p && f >= 0 && f < PQB_N_FIELDS) ? p->h_in[f] : nullptr;
}
extern "C" const double *pqb_panel_host_output(pqb_panel *p, int k) {
    return (p && k >= 0 && k < PQB_N_OUTPUTS) ? p->h_out[k] : nullptr;
}
extern "C" const uint8_t *pqb_panel_host_validity(pqb_panel *p, int k) {
    return (p && k >= 0 && k < PQB_N_OUTPUTS) ? reinterpret_cast<const uint8_t *>(p->h_bits[k]) : nullptr;
}
extern "C" const double *pqb_panel_device_field(const pqb_panel *p, int f) {
    return (p && f >= 0 && f < PQB_N_FIELDS) ? p->d_in[f] : nullptr;
}
extern "C" const double *pqb_panel_device_output(const pqb_panel *p, int k) {
    return (p && k >= 0 && k < PQB_N_OUTPUTS) ? p->d_out[k] : nullptr;
}
extern "C" const uint8_t *pqb_panel_device_validity(const pqb_panel *p, int k) {
    return (p && k >= 0 && k < PQB_N_OUTPUTS) ? reinterpret_cast<const uint8_t *>(p->d_bits[k]) : nullptr;
}

static inline bool bit_at(const uint8_t *bm, int64_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }

extern "C" int pqb_panel_set_column(pqb_panel *p, int64_t symbol, int field, const double *values,
                                    const uint8_t *validity, int64_t offset, int64_t len) {
    if (!p || !values) return fail(PQB_ERR_INVALID, "pqb_panel_set_column: NULL argument");
    if (!p->staging) return fail(PQB_ERR_INVALID, "pqb_panel_set_column: panel has no host staging");
    if (symbol < 0 || symbol >= p->n_symbols || field < 0 || field >= PQB_N_FIELDS || !p->h_in[field])
        return fail(PQB_ERR_INVALID, "pqb_panel_set_column: bad symbol/field");
    if (len != p->n_bars || offset < 0)
        return fail(PQB_ERR_INVALID, "pqb_panel_set_column: len %lld != n_bars %lld", (long long)len, (long long)p->n_bars);
    int64_t lead = 0;
    if (validity) {
        while (lead < len && !bit_at(validity, offset + lead)) ++lead;
        for (int64_t i = lead; i < len; ++i)
            if (!bit_at(validity, offset + i))
                return fail(PQB_ERR_UNSUPPORTED,
                            "pqb_panel_set_column: symbol %lld field %d has a null at %lld after its first valid bar "
                            "(interior/trailing nulls are not supported by the panel path)",
                            (long long)symbol, field, (long long)i);
    }
    double *dst = p->h_in[field] + (size_t)symbol * p->pitch;
    memcpy(dst, values + offset, (size_t)len * sizeof(double));
    for (int64_t i = 0; i < lead; ++i) dst[i] = 0.0;
    for (int64_t i = len; i < p->pitch; ++i) dst[i] = 0.0;
    if (lead > p->h_start[(size_t)symbol]) { p->h_start[(size_t)symbol] = (int32_t)lead; p->starts_nonzero = true; }
    return PQB_OK;
}

extern "C" int pqb_panel_set_starts(pqb_panel *p, const int32_t *starts) {
    if (!p || !starts) return fail(PQB_ERR_INVALID, "pqb_panel_set_starts: NULL argument");
    p->starts_nonzero = false;
    for (int64_t s = 0; s < p->n_symbols; ++s) {
        if (starts[s] < 0) return fail(PQB_ERR_INVALID, "pqb_panel_set_starts: negative start");
        p->h_start[(size_t)s] = std::min<int64_t>(starts[s], p->n_bars);
        if (starts[s]) p->starts_nonzero = true;
    }
    int rc = set_dev(p->e);
    if (rc) return rc;
    CU(cudaMemcpyAsync(p->d_start, p->h_start.data(), (size_t)p->n_symbols * sizeof(int), cudaMemcpyHostToDevice,
                       p->e->stream));
    CU(cudaStreamSynchronize(p->e->stream));
    return PQB_OK;
}

extern "C" int pqb_panel_upload(pqb_panel *p) {
    if (!p || !p->staging) return fail(PQB_ERR_INVALID, "pqb_panel_upload: no host staging");
    int rc = set_dev(p->e);
    if (rc) return rc;
    const size_t plane = (size_t)p->n_symbols * p->pitch * sizeof(double);
    for (int f = 0; f < PQB_N_FIELDS; ++f)
        if (p->d_in[f]) CU(cudaMemcpyAsync(p->d_in[f], p->h_in[f], plane, cudaMemcpyHostToDevice, p->e->stream));
    CU(cudaMemcpyAsync(p->d_start, p->h_start.data(), (size_t)p->n_symbols * sizeof(int), cudaMemcpyHostToDevice,
                       p->e->stream));
    return PQB_OK;
}

extern "C" int pqb_panel_download(pqb_panel *p) {
    if (!p || !p->staging) return fail(PQB_ERR_INVALID, "pqb_panel_download: no host staging");
    int rc = set_dev(p->e);
    if (rc) return rc;
    const size_t plane = (size_t)p->n_symbols * p->pitch * sizeof(double);
    const size_t bplane = (size_t)p->n_symbols * p->words_per_row * sizeof(uint32_t);
    for (int k = 0; k < PQB_N_OUTPUTS; ++k) {
        if (!p->d_out[k]) continue;
        CU(cudaMemcpyAsync(p->h_out[k], p->d_out[k], plane, cudaMemcpyDeviceToHost, p->e->stream));
        CU(cudaMemcpyAsync(p->h_bits[k], p->d_bits[k], bplane, cudaMemcpyDeviceToHost, p->e->stream));
    }
    return PQB_OK;
}

extern "C" int pqb_panel_sync(pqb_panel *p) {
    if (!p) return fail(PQB_ERR_INVALID, "pqb_panel_sync: NULL");
    int rc = set_dev(p->e);
    if (rc) return rc;
    CU(cudaStreamSynchronize(p->e->h2d));
    CU(cudaStreamSynchronize(p->e->stream));
    CU(cudaStreamSynchronize(p->e->d2h));
    return PQB_OK;
}

extern "C" int pqb_panel_get_output(pqb_panel *p, int64_t symbol, int output, double *values, uint8_t *validity,
                                    int64_t len) {
    if (!p || !values) return fail(PQB_ERR_INVALID, "pqb_panel_get_output: NULL argument");
    if (!p->staging || output < 0 || output >= PQB_N_OUTPUTS || !p->h_out[output] || symbol < 0 ||
        symbol >= p->n_symbols || len != p->n_bars)
        return fail(PQB_ERR_INVALID, "pqb_panel_get_output: bad symbol/output/len");
    memcpy(values, p->h_out[output] + (size_t)symbol * p->pitch, (size_t)len * sizeof(double));
    if (validity)
        memcpy(validity, reinterpret_cast<const uint8_t *>(p->h_bits[output] + (size_t)symbol * p->words_per_row),
               (size_t)((len + 7) / 8));
    return PQB_OK;
}

// ---------------------------------------------------------------------------------------
// parameters -> kernel arguments
// ---------------------------------------------------------------------------------------
static EmaK make_ema_k(int p, double alpha) {
    EmaK k{};
    k.alpha = alpha;
    k.p = p;
    k.pd = (double)p;
    const double a = 1.0 - alpha;
    double pw = a;
    for (int i = 0; i < 4; ++i) { k.pw[i] = pw; pw *= a; }
    double A = k.pw[3];
    for (int j = 0; j < 5; ++j) { k.A[j] = A; A *= A; }
    return k;
}
static inline double ema_alpha(int p) { return 2.0 / ((double)p + 1.0); }   // overlap.rs:669

struct Built {
    SuiteArgs a;
    int halo;          // 32 or 128
};

static int build_args(const pqb_panel *p, const pqb_suite_params *sp, Built *out) {
    SuiteArgs &A = out->a;
    memset(&A, 0, sizeof A);
    const int n_bars = (int)p->n_bars;
    const int NEVER = n_bars;                       // lead that makes a column all-null
    uint32_t ind = sp->indicators & PQB_IND_ALL;
    auto need_fields = [&](uint32_t mask, const char *what) -> int {
        if ((p->fields_mask & mask) != mask) return fail(PQB_ERR_INVALID, "%s needs panel fields 0x%x", what, mask);
        return PQB_OK;
    };
    const uint32_t C = 1u << PQB_CLOSE, H = 1u << PQB_HIGH, L = 1u << PQB_LOW, V = 1u << PQB_VOLUME;
    int rc;
    if ((ind & (PQB_IND_SMA | PQB_IND_EMA | PQB_IND_TEMA | PQB_IND_TRIMA | PQB_IND_BBANDS | PQB_IND_MACD | PQB_IND_RSI)) &&
        (rc = need_fields(C, "close-based indicators"))) return rc;
    if ((ind & (PQB_IND_TRANGE | PQB_IND_ATR | PQB_IND_NATR | PQB_IND_KDJ | PQB_IND_WILLR)) &&
        (rc = need_fields(C | H | L, "high/low/close indicators"))) return rc;
    if ((ind & PQB_IND_MIDPRICE) && (rc = need_fields(H | L, "midprice"))) return rc;
    if ((ind & PQB_IND_OBV) && (rc = need_fields(C | V, "obv"))) return rc;
    if ((ind & PQB_IND_AD) && (rc = need_fields(C | H | L | V, "ad"))) return rc;

    for (int f = 0; f < PQB_N_FIELDS; ++f) A.in[f] = p->d_in[f] ? p->d_in[f] : nullptr;
    // the TMA producer always copies all four planes of a tile: alias missing ones to a present one
    const double *any = nullptr;
    for (int f = 0; f < PQB_N_FIELDS; ++f) if (A.in[f]) { any = A.in[f]; break; }
    for (int f = 0; f < PQB_N_FIELDS; ++f) if (!A.in[f]) A.in[f] = any;
    for (int k = 0; k < PQB_N_OUTPUTS; ++k) { A.out[k] = nullptr; A.lead[k] = NEVER; }
    A.start = p->starts_nonzero ? p->d_start : nullptr;
    A.n_symbols = (int)p->n_symbols;
    A.n_bars = n_bars;
    A.pitch = (int)p->pitch;

    int max_sum_window = 1, max_ext_window = 1;
    auto neg = [&](int v, const char *name) -> int {
        return v < 0 ? fail(PQB_ERR_INVALID, "%s period %d is negative", name, v) : PQB_OK;
    };
    auto want = [&](int k) { return p->d_out[k] != nullptr; };
    auto bind = [&](int k, long long lead) {
        if (want(k)) { A.out[k] = p->d_out[k]; A.lead[k] = (int)std::min<long long>(lead, NEVER); }
    };
    // A period of 0 makes the reference return an all-null column (guards overlap.rs:663,874,...):
    // the output is bound with lead = NEVER and its group stays off.
    auto null_only = [&](int k) { if (want(k)) { A.out[k] = p->d_out[k]; A.lead[k] = NEVER; } };
    // (all-null columns are NaN-filled by run_suite: no group stores them)

    if (ind & PQB_IND_SMA) {
        if ((rc = neg(sp->sma_period, "sma"))) return rc;
        if (sp->sma_period == 0) { null_only(PQB_OUT_SMA); }
        else { A.groups |= G_SMA; A.sma_p = sp->sma_period; A.inv_sma = 1.0 / (double)sp->sma_period;
               bind(PQB_OUT_SMA, sp->sma_period - 1); max_sum_window = std::max(max_sum_window, sp->sma_period); }
    }
    if (ind & PQB_IND_TEMA) {
        if ((rc = neg(sp->tema_period, "tema"))) return rc;
        const int tp = sp->tema_period;
        // guard overlap.rs:1180: n < 3p-2 -> all null (equivalent to lead >= n)
        if (tp == 0) { null_only(PQB_OUT_TEMA); }
        else { A.groups |= G_TEMA; A.k_tema = make_ema_k(tp, ema_alpha(tp)); bind(PQB_OUT_TEMA, 3ll * tp - 3); }
    }
    if (ind & PQB_IND_EMA) {
        if ((rc = neg(sp->ema_period, "ema"))) return rc;
        const int ep = sp->ema_period;
        if (ep == 0) { null_only(PQB_OUT_EMA); }
        else { A.groups |= G_EMA; A.k_ema = make_ema_k(ep, ema_alpha(ep)); bind(PQB_OUT_EMA, ep - 1);
               A.ema_shares_tema = (A.groups & G_TEMA) && sp->tema_period == ep; }
    }
    if (ind & PQB_IND_TRIMA) {
        if ((rc = neg(sp->trima_period, "trima"))) return rc;
        const int tp = sp->trima_period;
        int n1, n2;
        if (tp % 2 == 1) { n1 = tp / 2 + 1; n2 = n1; } else { n1 = tp / 2; n2 = n1 + 1; }   // overlap.rs:1314-1323
        if (n1 == 0) { null_only(PQB_OUT_TRIMA); }
        else { A.groups |= G_TRIMA; A.tri_n1 = n1; A.tri_n2 = n2; A.inv_tri1 = 1.0 / (double)n1; A.inv_tri2 = 1.0 / (double)n2;
               bind(PQB_OUT_TRIMA, (long long)n1 + n2 - 2); max_sum_window = std::max(max_sum_window, std::max(n1, n2)); }
    }
    if (ind & PQB_IND_BBANDS) {
        if ((rc = neg(sp->bbands_period, "bbands"))) return rc;
        const int bp = sp->bbands_period;
        if (bp == 0) { for (int k = 4; k <= 6; ++k) { null_only(k); } }
        else { A.groups |= G_BB; A.bb_p = bp; A.bb_pd = (double)bp; A.inv_bb = 1.0 / (double)bp;
               A.bb_up = sp->bbands_nbdevup; A.bb_dn = sp->bbands_nbdevdn;
               for (int k = 4; k <= 6; ++k) bind(k, bp - 1);
               max_sum_window = std::max(max_sum_window, bp); }
    }
    if (ind & PQB_IND_MACD) {
        const int f = sp->macd_fast, s = sp->macd_slow, g = sp->macd_signal;
        if ((rc = neg(f, "macd fast")) || (rc = neg(s, "macd slow")) || (rc = neg(g, "macd signal"))) return rc;
        if (f == 0 || s == 0 || g == 0)
            return fail(PQB_ERR_UNSUPPORTED, "macd with a zero period (reference yields partial nulls) is not built");
        A.groups |= G_MACD;
        A.k_macd_f = make_ema_k(f, ema_alpha(f)); A.k_macd_s = make_ema_k(s, ema_alpha(s)); A.k_macd_g = make_ema_k(g, ema_alpha(g));
        A.macd_dif_lead = std::max(f, s) - 1;
        bind(PQB_OUT_MACD, A.macd_dif_lead); bind(PQB_OUT_MACD_SIGNAL, g - 1);
        bind(PQB_OUT_MACD_HIST, std::max(A.macd_dif_lead, g - 1));
    }
    if (ind & PQB_IND_RSI) {
        if ((rc = neg(sp->rsi_period, "rsi"))) return rc;
        const int rp = sp->rsi_period;
        if (rp == 0) { null_only(PQB_OUT_RSI); }
        else { A.groups |= G_RSI; A.k_rsi = make_ema_k(rp, 1.0 / (double)rp); bind(PQB_OUT_RSI, rp - 1); }   // D1
    }
    if (ind & PQB_IND_TRANGE) { A.groups |= G_TRANGE; bind(PQB_OUT_TRANGE, 1); }
    if (ind & PQB_IND_ATR) {
        if (sp->atr_period <= 0) return fail(PQB_ERR_INVALID, "atr period %d: 2p-1 underflows in the reference", sp->atr_period);
        const int ep = 2 * sp->atr_period - 1;                                               // volatility.rs:30
        A.groups |= G_ATR; A.k_atr = make_ema_k(ep, ema_alpha(ep)); bind(PQB_OUT_ATR, (long long)ep);
    }
    if (ind & PQB_IND_NATR) {
        if (sp->natr_period <= 0) return fail(PQB_ERR_INVALID, "natr period %d: 2p-1 underflows in the reference", sp->natr_period);
        const int ep = 2 * sp->natr_period - 1;
        A.groups |= G_NATR; A.k_natr = make_ema_k(ep, ema_alpha(ep)); bind(PQB_OUT_NATR, (long long)ep);
        A.natr_shares_atr = (A.groups & G_ATR) && sp->atr_period == sp->natr_period;
    }
    if (ind & PQB_IND_OBV) { A.groups |= G_OBV; bind(PQB_OUT_OBV, 1); }
    if (ind & PQB_IND_AD) { A.groups |= G_AD; bind(PQB_OUT_AD, 0); }
    if (ind & PQB_IND_KDJ) {
        const int k = sp->kdj_fastk, sk = sp->kdj_slowk, sd = sp->kdj_slowd;
        if ((rc = neg(k, "kdj fastk")) || (rc = neg(sk, "kdj slowk")) || (rc = neg(sd, "kdj slowd"))) return rc;
        if (k == 0 || sk == 0 || sd == 0) {
            for (int q = 16; q <= 18; ++q) { null_only(q); }
            if (k != 0 && sk != 0 && sd == 0)
                return fail(PQB_ERR_UNSUPPORTED, "kdj with slowd_period 0 (K valid, D null) is not built");
        } else {
            A.groups |= G_KDJ; A.kdj_k = k; A.kdj_sk = sk; A.kdj_sd = sd; A.inv_sk = 1.0 / (double)sk; A.inv_sd = 1.0 / (double)sd;
            bind(PQB_OUT_KDJ_K, (long long)k + sk - 2); bind(PQB_OUT_KDJ_D, (long long)k + sk + sd - 3);
            bind(PQB_OUT_KDJ_J, (long long)k + sk + sd - 3);
            max_ext_window = std::max(max_ext_window, k); max_sum_window = std::max(max_sum_window, std::max(sk, sd));
        }
    }
    if (ind & PQB_IND_WILLR) {
        if ((rc = neg(sp->willr_period, "willr"))) return rc;
        if (sp->willr_period == 0) { null_only(PQB_OUT_WILLR); }
        else { A.groups |= G_WILLR; A.willr_p = sp->willr_period; bind(PQB_OUT_WILLR, sp->willr_period - 1);
               max_ext_window = std::max(max_ext_window, sp->willr_period); }
    }
    if (ind & PQB_IND_MIDPRICE) {
        if (sp->midprice_period <= 0)
            return fail(PQB_ERR_UNSUPPORTED, "midprice period %d (reference: never-expiring deque) is not built", sp->midprice_period);
        A.groups |= G_MIDPRICE; A.mid_p = sp->midprice_period; bind(PQB_OUT_MIDPRICE, 0);
        max_ext_window = std::max(max_ext_window, sp->midprice_period);
    }
    {
        bool all_bound = true;
        int L = 0;
        for (int k = 0; k < PQB_N_OUTPUTS; ++k) { all_bound &= A.out[k] != nullptr; L = std::max(L, A.lead[k]); }
        A.steady_ok = (A.groups == G_ALL) && all_bound && A.ema_shares_tema && A.natr_shares_atr;
        if (A.steady_ok && A.sma_p == 30 && A.tri_n1 == 15 && A.tri_n2 == 16 && A.bb_p == 20 && A.kdj_k == 9 &&
            A.kdj_sk == 3 && A.kdj_sd == 3 && A.willr_p == 14 && A.mid_p == 14)
            A.steady_ok = 2;        // window periods == the reference's Python defaults: baked-in path
        A.steady_lead = L + 1;
    }
    if (max_ext_window > 32)
        return fail(PQB_ERR_UNSUPPORTED, "rolling max/min window %d > 32 is not built yet", max_ext_window);
    if (max_sum_window > 128)
        return fail(PQB_ERR_UNSUPPORTED, "windowed-sum period %d > 128 is not built yet", max_sum_window);
    out->halo = (max_sum_window > 32) ? 128 : 32;
    return PQB_OK;
}

// NaN fill for all-null columns (period 0)
__global__ void nan_fill_kernel(double *p, size_t n) {
    const double nn = __longlong_as_double(0x7ff8000000000000LL);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = nn;
}

static int launch_suite(pqb_engine *e, const Built &b, uint32_t *const *bits, int words_per_row,
                        cudaEvent_t ev_after_fused, int *launches) {
    int n_launch = 0;
    if (b.a.groups) {
        const int per_sm = (b.halo == 32) ? e->ctas_per_sm32 : e->ctas_per_sm128;
        const long long warps_needed = b.a.n_symbols;
        int grid = e->sm_count * per_sm;
        grid = (int)std::min<long long>(grid, (warps_needed + kWarpsPerCta - 1) / kWarpsPerCta);
        if (b.halo == 32)
            suite_fused_kernel<32><<<grid, PQB_CTA_THREADS, kWarpsPerCta * WarpSmem<32>::BYTES, e->stream>>>(b.a);
        else
            suite_fused_kernel<128><<<grid, PQB_CTA_THREADS, kWarpsPerCta * WarpSmem<128>::BYTES, e->stream>>>(b.a);
        CU(cudaGetLastError());
        ++n_launch;
    }
    if (ev_after_fused) CU(cudaEventRecord(ev_after_fused, e->stream));
    ValidityArgs V{};
    bool any = false;
    for (int k = 0; k < PQB_N_OUTPUTS; ++k) {
        V.bits[k] = b.a.out[k] ? bits[k] : nullptr;
        V.lead[k] = b.a.lead[k];
        any |= V.bits[k] != nullptr;
    }
    V.start = b.a.start;
    V.n_symbols = b.a.n_symbols;
    V.n_bars = b.a.n_bars;
    V.words_per_row = words_per_row;
    if (any) {
        const long long total = (long long)V.n_symbols * V.words_per_row;
        const int grid = (int)std::min<long long>((total + 255) / 256, (long long)e->sm_count * 8);
        validity_kernel<<<grid, 256, 0, e->stream>>>(V);
        CU(cudaGetLastError());
        ++n_launch;
    }
    if (launches) *launches = n_launch;
    return PQB_OK;
}

// which outputs does an enabled group actually store?  (used for the period-0 NaN fill)
static uint32_t outputs_of_groups(unsigned g) {
    uint32_t m = 0;
    if (g & G_SMA) m |= 1u << 0;
    if (g & G_EMA) m |= 1u << 1;
    if (g & G_TEMA) m |= 1u << 2;
    if (g & G_TRIMA) m |= 1u << 3;
    if (g & G_BB) m |= 7u << 4;
    if (g & G_MACD) m |= 7u << 7;
    if (g & G_RSI) m |= 1u << 10;
    if (g & G_TRANGE) m |= 1u << 11;
    if (g & G_ATR) m |= 1u << 12;
    if (g & G_NATR) m |= 1u << 13;
    if (g & G_OBV) m |= 1u << 14;
    if (g & G_AD) m |= 1u << 15;
    if (g & G_KDJ) m |= 7u << 16;
    if (g & G_WILLR) m |= 1u << 19;
    if (g & G_MIDPRICE) m |= 1u << 20;
    return m;
}

static int run_suite(pqb_panel *p, const pqb_suite_params *sp, cudaEvent_t ev_after_fused, int *launches) {
    if (!p || !sp) return fail(PQB_ERR_INVALID, "pqb_suite_run: NULL argument");
    int rc = set_dev(p->e);
    if (rc) return rc;
    Built b;
    if ((rc = build_args(p, sp, &b))) return rc;
    int extra = 0;
    const uint32_t stored = outputs_of_groups(b.a.groups);
    for (int k = 0; k < PQB_N_OUTPUTS; ++k) {
        if (b.a.out[k] && !(stored >> k & 1)) {
            const size_t n = (size_t)p->n_symbols * p->pitch;
            nan_fill_kernel<<<p->e->sm_count * 4, 256, 0, p->e->stream>>>(b.a.out[k], n);
            CU(cudaGetLastError());
            ++extra;
        }
    }
    int nl = 0;
    rc = launch_suite(p->e, b, p->d_bits, (int)p->words_per_row, ev_after_fused, &nl);
    if (launches) *launches = nl + extra;
    return rc;
}

extern "C" int pqb_suite_run(pqb_panel *p, const pqb_suite_params *sp) { return run_suite(p, sp, nullptr, nullptr); }

// ---------------------------------------------------------------------------------------
// end-to-end host path: chunked upload -> suite -> download over pinned staging
// ---------------------------------------------------------------------------------------
extern "C" int pqb_suite_run_host(pqb_panel *p, const pqb_suite_params *sp, int64_t chunk_symbols) {
    if (!p || !sp) return fail(PQB_ERR_INVALID, "pqb_suite_run_host: NULL argument");
    if (!p->staging) return fail(PQB_ERR_INVALID, "pqb_suite_run_host: panel has no host staging");
    int rc = set_dev(p->e);
    if (rc) return rc;
    pqb_engine *e = p->e;
    Built full;
    if ((rc = build_args(p, sp, &full))) return rc;
    if (chunk_symbols <= 0) chunk_symbols = std::max<int64_t>(1, (int64_t)e->sm_count * 8);
    const int64_t n_chunks = (p->n_symbols + chunk_symbols - 1) / chunk_symbols;
    const uint32_t stored = outputs_of_groups(full.a.groups);
    std::vector<cudaEvent_t> up((size_t)n_chunks), done((size_t)n_chunks);
    for (auto &x : up) CU(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    for (auto &x : done) CU(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    CU(cudaMemcpyAsync(p->d_start, p->h_start.data(), (size_t)p->n_symbols * sizeof(int), cudaMemcpyHostToDevice, e->h2d));
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t s0 = c * chunk_symbols, ns = std::min(chunk_symbols, p->n_symbols - s0);
        const size_t off = (size_t)s0 * p->pitch, bytes = (size_t)ns * p->pitch * sizeof(double);
        for (int f = 0; f < PQB_N_FIELDS; ++f)
            if (p->d_in[f]) CU(cudaMemcpyAsync(p->d_in[f] + off, p->h_in[f] + off, bytes, cudaMemcpyHostToDevice, e->h2d));
        CU(cudaEventRecord(up[(size_t)c], e->h2d));
        CU(cudaStreamWaitEvent(e->stream, up[(size_t)c], 0));
        // a view of the panel restricted to this chunk of symbols
        Built b = full;
        for (int f = 0; f < PQB_N_FIELDS; ++f) b.a.in[f] = full.a.in[f] + off;
        for (int k = 0; k < PQB_N_OUTPUTS; ++k) if (full.a.out[k]) b.a.out[k] = full.a.out[k] + off;
        if (full.a.start) b.a.start = full.a.start + s0;
        b.a.n_symbols = (int)ns;
        uint32_t *bits[PQB_N_OUTPUTS];
        for (int k = 0; k < PQB_N_OUTPUTS; ++k) bits[k] = p->d_bits[k] ? p->d_bits[k] + (size_t)s0 * p->words_per_row : nullptr;
        for (int k = 0; k < PQB_N_OUTPUTS; ++k) {
            if (b.a.out[k] && !(stored >> k & 1)) {
                nan_fill_kernel<<<e->sm_count * 4, 256, 0, e->stream>>>(b.a.out[k], (size_t)ns * p->pitch);
                CU(cudaGetLastError());
            }
        }
        if ((rc = launch_suite(e, b, bits, (int)p->words_per_row, nullptr, nullptr))) return rc;
        CU(cudaEventRecord(done[(size_t)c], e->stream));
        CU(cudaStreamWaitEvent(e->d2h, done[(size_t)c], 0));
        const size_t boff = (size_t)s0 * p->words_per_row, bbytes = (size_t)ns * p->words_per_row * sizeof(uint32_t);
        for (int k = 0; k < PQB_N_OUTPUTS; ++k) {
            if (!full.a.out[k]) continue;
            CU(cudaMemcpyAsync(p->h_out[k] + off, p->d_out[k] + off, bytes, cudaMemcpyDeviceToHost, e->d2h));
            CU(cudaMemcpyAsync(p->h_bits[k] + boff, p->d_bits[k] + boff, bbytes, cudaMemcpyDeviceToHost, e->d2h));
        }
    }
    CU(cudaStreamSynchronize(e->d2h));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaStreamSynchronize(e->h2d));
    for (auto &x : up) cudaEventDestroy(x);
    for (auto &x : done) cudaEventDestroy(x);
    return PQB_OK;
}

// ---------------------------------------------------------------------------------------
// measurement helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
// approx N(0,1): Irwin-Hall sum of 12 uniforms built from 6 hashes (2 x 32-bit halves each)
__device__ __forceinline__ double gauss(uint64_t key) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const uint64_t h = mix64(key * 6 + i);
        s += (double)(uint32_t)h * (1.0 / 4294967296.0) + (double)(uint32_t)(h >> 32) * (1.0 / 4294967296.0);
    }
    return s - 6.0;
}

// 32 symbols per warp; each lane walks its symbol 32 bars at a time, the warp transposes through
// shared memory so global writes are 256 B contiguous per row.
__global__ void __launch_bounds__(32) synth_kernel(double *c, double *h, double *l, double *v, int n_symbols,
                                                   int n_bars, int pitch, uint64_t seed, double sigma) {
    __shared__ double tc[32][33], th[32][33], tlo[32][33], tv[32][33];
    const int lane = threadIdx.x;
    const int s0 = blockIdx.x * 32;
    const int s = s0 + lane;
    double close = 100.0;
    for (int t0 = 0; t0 < pitch; t0 += 32) {
        for (int j = 0; j < 32; ++j) {
            const int t = t0 + j;
            const uint64_t key = (seed + (uint64_t)s) * 0x100000001b3ULL + (uint64_t)t * 4;
            const double open = close;
            close = open * exp(sigma * gauss(key));
            const double hi = fmax(open, close) * (1.0 + fabs(0.5 * sigma * gauss(key + 1)));
            const double lo = fmin(open, close) * (1.0 - fabs(0.5 * sigma * gauss(key + 2)));
            const double vol = rint(exp(13.0 + gauss(key + 3)));
            const bool live = t < n_bars;
            tc[lane][j] = live ? close : 0.0;
            th[lane][j] = live ? hi : 0.0;
            tlo[lane][j] = live ? lo : 0.0;
            tv[lane][j] = live ? vol : 0.0;
        }
        __syncwarp();
        for (int r = 0; r < 32; ++r) {
            const int sr = s0 + r;
            if (sr < n_symbols && t0 + lane < pitch) {
                const size_t o = (size_t)sr * pitch + t0 + lane;
                if (c) c[o] = tc[r][lane];
                if (h) h[o] = th[r][lane];
                if (l) l[o] = tlo[r][lane];
                if (v) v[o] = tv[r][lane];
            }
        }
        __syncwarp();
    }
}

extern "C" int pqb_panel_fill_synthetic(pqb_panel *p, uint64_t seed, double sigma, int to_host) {
    if (!p) return fail(PQB_ERR_INVALID, "pqb_panel_fill_synthetic: NULL");
    int rc = set_dev(p->e);
    if (rc) return rc;
    const int grid = (int)((p->n_symbols + 31) / 32);
    synth_kernel<<<grid, 32, 0, p->e->stream>>>(p->d_in[PQB_CLOSE], p->d_in[PQB_HIGH], p->d_in[PQB_LOW],
                                                 p->d_in[PQB_VOLUME], (int)p->n_symbols, (int)p->n_bars, (int)p->pitch,
                                                 seed, sigma);
    CU(cudaGetLastError());
    if (to_host && p->staging) {
        const size_t plane = (size_t)p->n_symbols * p->pitch * sizeof(double);
        for (int f = 0; f < PQB_N_FIELDS; ++f)
            if (p->d_in[f]) CU(cudaMemcpyAsync(p->h_in[f], p->d_in[f], plane, cudaMemcpyDeviceToHost, p->e->stream));
    }
    CU(cudaStreamSynchronize(p->e->stream));
    return PQB_OK;
}

extern "C" int pqb_suite_time(pqb_panel *p, const pqb_suite_params *sp, int warmup, int iters, float *ms_total,
                              float *ms_fused, int *launches_per_step) {
    if (!p || !sp || iters <= 0) return fail(PQB_ERR_INVALID, "pqb_suite_time: bad argument");
    int rc = set_dev(p->e);
    if (rc) return rc;
    cudaStream_t st = p->e->stream;
    int nl = 0;
    for (int i = 0; i < warmup; ++i)
        if ((rc = run_suite(p, sp, nullptr, &nl))) return rc;
    CU(cudaStreamSynchronize(st));
    std::vector<cudaEvent_t> b((size_t)iters), a((size_t)iters);
    for (auto &x : b) CU(cudaEventCreate(&x));
    for (auto &x : a) CU(cudaEventCreate(&x));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) {
        CU(cudaEventRecord(b[(size_t)i], st));
        if ((rc = run_suite(p, sp, a[(size_t)i], &nl))) return rc;
    }
    CU(cudaEventRecord(e1, st));
    CU(cudaStreamSynchronize(st));
    float tot = 0.f, fused = 0.f;
    CU(cudaEventElapsedTime(&tot, e0, e1));
    for (int i = 0; i < iters; ++i) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, b[(size_t)i], a[(size_t)i]));
        fused += t;
    }
    for (auto &x : b) cudaEventDestroy(x);
    for (auto &x : a) cudaEventDestroy(x);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_total) *ms_total = tot;
    if (ms_fused) *ms_fused = fused;
    if (launches_per_step) *launches_per_step = nl;
    return PQB_OK;
}

extern "C" int pqb_suite_time_host(pqb_panel *p, const pqb_suite_params *sp, int64_t chunk_symbols, int warmup,
                                   int iters, float *ms_total) {
    if (!p || !sp || iters <= 0) return fail(PQB_ERR_INVALID, "pqb_suite_time_host: bad argument");
    int rc = set_dev(p->e);
    if (rc) return rc;
    for (int i = 0; i < warmup; ++i)
        if ((rc = pqb_suite_run_host(p, sp, chunk_symbols))) return rc;
    // the whole pipeline spans three streams; bracket it on the host with device-wide syncs and
    // CUDA events on a stream that joins all three
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaDeviceSynchronize());
    CU(cudaEventRecord(e0, p->e->h2d));
    for (int i = 0; i < iters; ++i)
        if ((rc = pqb_suite_run_host(p, sp, chunk_symbols))) return rc;
    CU(cudaEventRecord(e1, p->e->d2h));
    CU(cudaEventSynchronize(e1));
    float tot = 0.f;
    CU(cudaEventElapsedTime(&tot, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_total) *ms_total = tot;
    return PQB_OK;
}

__global__ void flush_kernel(uint4 *p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4((unsigned)i, 1u, 2u, 3u);
}
extern "C" int pqb_flush_l2(pqb_engine *e) {
    if (!e) return fail(PQB_ERR_INVALID, "pqb_flush_l2: NULL");
    int rc = set_dev(e);
    if (rc) return rc;
    if (!e->flush_buf) {
        e->flush_bytes = 256ull << 20;   // 2x the 126 MB L2
        CU(cudaMalloc(&e->flush_buf, e->flush_bytes));
    }
    flush_kernel<<<e->sm_count * 4, 256, 0, e->stream>>>((uint4 *)e->flush_buf, e->flush_bytes / 16);
    CU(cudaGetLastError());
    return PQB_OK;
}

// ---------------------------------------------------------------------------------------
// single-column entry points: one reference plugin call on one column
// ---------------------------------------------------------------------------------------
struct ColCheck { int64_t lead; bool any_null; bool interior; };
static ColCheck scan_col(const pqb_col *c) {
    ColCheck r{0, false, false};
    if (!c->validity) return r;
    while (r.lead < c->len && !bit_at(c->validity, c->offset + r.lead)) ++r.lead;
    r.any_null = r.lead > 0;
    for (int64_t i = r.lead; i < c->len; ++i)
        if (!bit_at(c->validity, c->offset + i)) { r.any_null = true; r.interior = true; break; }
    return r;
}

enum NullPolicy { NP_SHIFT, NP_ERR };   // overlap/volatility/volume functions skip nulls; momentum.rs errors

static int run_single(pqb_engine *e, const pqb_col *const *cols, const int *fields, int n_cols, NullPolicy np,
                      const pqb_suite_params *sp, const int *outs, pqb_out_col *const *dst, int n_out) {
    if (!e) return fail(PQB_ERR_INVALID, "NULL engine");
    for (int i = 0; i < n_cols; ++i)
        if (!cols[i] || (!cols[i]->values && cols[i]->len > 0) || cols[i]->len < 0 || cols[i]->offset < 0)
            return fail(PQB_ERR_INVALID, "bad input column %d", i);
    for (int i = 0; i < n_out; ++i)
        if (!dst[i] || (!dst[i]->values && cols[0]->len > 0) || (!dst[i]->validity && cols[0]->len > 0))
            return fail(PQB_ERR_INVALID, "bad output column %d", i);
    const int64_t n = cols[0]->len;
    for (int i = 1; i < n_cols; ++i)
        if (cols[i]->len != n) return fail(PQB_ERR_INVALID, "input columns differ in length");
    int64_t lead = -1;
    for (int i = 0; i < n_cols; ++i) {
        ColCheck cc = scan_col(cols[i]);
        if (cc.any_null && np == NP_ERR)
            return fail(PQB_ERR_NULLS, "chunked array is not contiguous (input %d has nulls; reference: cont_slice()?)", i);
        if (cc.interior)
            return fail(PQB_ERR_UNSUPPORTED, "input %d has nulls after its first valid value (not built)", i);
        if (lead >= 0 && cc.lead != lead)
            return fail(PQB_ERR_UNSUPPORTED, "inputs start at different rows (%lld vs %lld; not built)", (long long)lead,
                        (long long)cc.lead);
        lead = cc.lead;
    }
    if (n == 0) return PQB_OK;
    // device required from here on
    int rc = set_dev(e);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->scratch || e->scratch_bars != n) {
        if (e->scratch) { pqb_panel_destroy(e->scratch); e->scratch = nullptr; }
        rc = pqb_panel_create(e, 1, n, (1u << PQB_N_FIELDS) - 1, (1u << PQB_N_OUTPUTS) - 1, 1, &e->scratch);
        if (rc) return rc;
        e->scratch_bars = n;
    }
    pqb_panel *p = e->scratch;
    for (int f = 0; f < PQB_N_FIELDS; ++f) memset(p->h_in[f], 0, (size_t)p->pitch * sizeof(double));
    p->h_start[0] = 0;
    p->starts_nonzero = false;
    for (int i = 0; i < n_cols; ++i)
        if ((rc = pqb_panel_set_column(p, 0, fields[i], cols[i]->values, cols[i]->validity, cols[i]->offset, n))) return rc;
    if ((rc = pqb_panel_upload(p))) return rc;
    if ((rc = pqb_suite_run(p, sp))) return rc;
    if ((rc = pqb_panel_download(p))) return rc;
    if ((rc = pqb_panel_sync(p))) return rc;
    for (int i = 0; i < n_out; ++i)
        if ((rc = pqb_panel_get_output(p, 0, outs[i], dst[i]->values, dst[i]->validity, n))) return rc;
    return PQB_OK;
}

static pqb_suite_params only(uint32_t ind) {
    pqb_suite_params sp;
    pqb_suite_params_default(&sp);
    sp.indicators = ind;
    return sp;
}

extern "C" int pqb_sma(pqb_engine *e, const pqb_col *real, int32_t tp, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_SMA); sp.sma_period = tp;
    const pqb_col *c[] = {real}; const int f[] = {PQB_CLOSE}; const int o[] = {PQB_OUT_SMA}; pqb_out_col *d[] = {out};
    return run_single(e, c, f, 1, NP_SHIFT, &sp, o, d, 1);
}
extern "C" int pqb_ema(pqb_engine *e, const pqb_col *real, int32_t tp, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_EMA); sp.ema_period = tp;
    const pqb_col *c[] = {real}; const int f[] = {PQB_CLOSE}; const int o[] = {PQB_OUT_EMA}; pqb_out_col *d[] = {out};
    return run_single(e, c, f, 1, NP_SHIFT, &sp, o, d, 1);
}
extern "C" int pqb_tema(pqb_engine *e, const pqb_col *real, int32_t tp, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_TEMA); sp.tema_period = tp;
    const pqb_col *c[] = {real}; const int f[] = {PQB_CLOSE}; const int o[] = {PQB_OUT_TEMA}; pqb_out_col *d[] = {out};
    return run_single(e, c, f, 1, NP_SHIFT, &sp, o, d, 1);
}
extern "C" int pqb_trima(pqb_engine *e, const pqb_col *real, int32_t tp, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_TRIMA); sp.trima_period = tp;
    const pqb_col *c[] = {real}; const int f[] = {PQB_CLOSE}; const int o[] = {PQB_OUT_TRIMA}; pqb_out_col *d[] = {out};
    return run_single(e, c, f, 1, NP_SHIFT, &sp, o, d, 1);
}
extern "C" int pqb_ma(pqb_engine *e, const pqb_col *real, int32_t tp, int32_t matype, pqb_out_col *out) {
    switch (matype) {                       // calc_ma overlap.rs:857-869
        case 1: return pqb_ema(e, real, tp, out);
        case 4: return pqb_tema(e, real, tp, out);
        case 5: return pqb_trima(e, real, tp, out);
        case 2: case 3: case 6: case 8:
            return fail(PQB_ERR_UNSUPPORTED, "matype %d (WMA/DEMA/KAMA/T3: defective in the reference, SURVEY 8a) is not built", matype);
        default: return pqb_sma(e, real, tp, out);
    }
}
extern "C" int pqb_bbands(pqb_engine *e, const pqb_col *real, int32_t tp, double up, double dn, pqb_out_col *u,
                          pqb_out_col *m, pqb_out_col *l) {
    pqb_suite_params sp = only(PQB_IND_BBANDS); sp.bbands_period = tp; sp.bbands_nbdevup = up; sp.bbands_nbdevdn = dn;
    const pqb_col *c[] = {real}; const int f[] = {PQB_CLOSE};
    const int o[] = {PQB_OUT_BB_UPPER, PQB_OUT_BB_MIDDLE, PQB_OUT_BB_LOWER}; pqb_out_col *d[] = {u, m, l};
    return run_single(e, c, f, 1, NP_SHIFT, &sp, o, d, 3);
}
extern "C" int pqb_macd(pqb_engine *e, const pqb_col *real, int32_t fp, int32_t slp, int32_t sgp, pqb_out_col *m,
                        pqb_out_col *s, pqb_out_col *h) {
    pqb_suite_params sp = only(PQB_IND_MACD); sp.macd_fast = fp; sp.macd_slow = slp; sp.macd_signal = sgp;
    const pqb_col *c[] = {real}; const int f[] = {PQB_CLOSE};
    const int o[] = {PQB_OUT_MACD, PQB_OUT_MACD_SIGNAL, PQB_OUT_MACD_HIST}; pqb_out_col *d[] = {m, s, h};
    return run_single(e, c, f, 1, NP_ERR, &sp, o, d, 3);
}
extern "C" int pqb_rsi(pqb_engine *e, const pqb_col *real, int32_t tp, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_RSI); sp.rsi_period = tp;
    const pqb_col *c[] = {real}; const int f[] = {PQB_CLOSE}; const int o[] = {PQB_OUT_RSI}; pqb_out_col *d[] = {out};
    return run_single(e, c, f, 1, NP_ERR, &sp, o, d, 1);
}
extern "C" int pqb_trange(pqb_engine *e, const pqb_col *h, const pqb_col *l, const pqb_col *cl, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_TRANGE);
    const pqb_col *c[] = {h, l, cl}; const int f[] = {PQB_HIGH, PQB_LOW, PQB_CLOSE}; const int o[] = {PQB_OUT_TRANGE};
    pqb_out_col *d[] = {out};
    return run_single(e, c, f, 3, NP_SHIFT, &sp, o, d, 1);
}
extern "C" int pqb_atr(pqb_engine *e, const pqb_col *h, const pqb_col *l, const pqb_col *cl, int32_t tp, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_ATR); sp.atr_period = tp;
    const pqb_col *c[] = {h, l, cl}; const int f[] = {PQB_HIGH, PQB_LOW, PQB_CLOSE}; const int o[] = {PQB_OUT_ATR};
    pqb_out_col *d[] = {out};
    return run_single(e, c, f, 3, NP_SHIFT, &sp, o, d, 1);
}
extern "C" int pqb_natr(pqb_engine *e, const pqb_col *h, const pqb_col *l, const pqb_col *cl, int32_t tp, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_NATR); sp.natr_period = tp;
    const pqb_col *c[] = {h, l, cl}; const int f[] = {PQB_HIGH, PQB_LOW, PQB_CLOSE}; const int o[] = {PQB_OUT_NATR};
    pqb_out_col *d[] = {out};
    return run_single(e, c, f, 3, NP_SHIFT, &sp, o, d, 1);
}
extern "C" int pqb_obv(pqb_engine *e, const pqb_col *cl, const pqb_col *v, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_OBV);
    const pqb_col *c[] = {cl, v}; const int f[] = {PQB_CLOSE, PQB_VOLUME}; const int o[] = {PQB_OUT_OBV};
    pqb_out_col *d[] = {out};
    return run_single(e, c, f, 2, NP_SHIFT, &sp, o, d, 1);
}
extern "C" int pqb_ad(pqb_engine *e, const pqb_col *h, const pqb_col *l, const pqb_col *cl, const pqb_col *v, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_AD);
    const pqb_col *c[] = {h, l, cl, v}; const int f[] = {PQB_HIGH, PQB_LOW, PQB_CLOSE, PQB_VOLUME};
    const int o[] = {PQB_OUT_AD}; pqb_out_col *d[] = {out};
    return run_single(e, c, f, 4, NP_SHIFT, &sp, o, d, 1);
}
extern "C" int pqb_stoch(pqb_engine *e, const pqb_col *h, const pqb_col *l, const pqb_col *cl, int32_t fk, int32_t sk,
                         int32_t sd, pqb_out_col *slowk, pqb_out_col *slowd) {
    pqb_suite_params sp = only(PQB_IND_KDJ); sp.kdj_fastk = fk; sp.kdj_slowk = sk; sp.kdj_slowd = sd;
    const pqb_col *c[] = {h, l, cl}; const int f[] = {PQB_HIGH, PQB_LOW, PQB_CLOSE};
    const int o[] = {PQB_OUT_KDJ_K, PQB_OUT_KDJ_D}; pqb_out_col *d[] = {slowk, slowd};
    return run_single(e, c, f, 3, NP_SHIFT, &sp, o, d, 2);
}
extern "C" int pqb_kdj(pqb_engine *e, const pqb_col *h, const pqb_col *l, const pqb_col *cl, int32_t fk, int32_t kp,
                       int32_t dp, pqb_out_col *k, pqb_out_col *dd, pqb_out_col *j) {
    pqb_suite_params sp = only(PQB_IND_KDJ); sp.kdj_fastk = fk; sp.kdj_slowk = kp; sp.kdj_slowd = dp;
    const pqb_col *c[] = {h, l, cl}; const int f[] = {PQB_HIGH, PQB_LOW, PQB_CLOSE};
    const int o[] = {PQB_OUT_KDJ_K, PQB_OUT_KDJ_D, PQB_OUT_KDJ_J}; pqb_out_col *d[] = {k, dd, j};
    return run_single(e, c, f, 3, NP_SHIFT, &sp, o, d, 3);
}
extern "C" int pqb_willr(pqb_engine *e, const pqb_col *h, const pqb_col *l, const pqb_col *cl, int32_t tp, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_WILLR); sp.willr_period = tp;
    const pqb_col *c[] = {h, l, cl}; const int f[] = {PQB_HIGH, PQB_LOW, PQB_CLOSE}; const int o[] = {PQB_OUT_WILLR};
    pqb_out_col *d[] = {out};
    return run_single(e, c, f, 3, NP_ERR, &sp, o, d, 1);
}
extern "C" int pqb_midprice(pqb_engine *e, const pqb_col *h, const pqb_col *l, int32_t tp, pqb_out_col *out) {
    pqb_suite_params sp = only(PQB_IND_MIDPRICE); sp.midprice_period = tp;
    const pqb_col *c[] = {h, l}; const int f[] = {PQB_HIGH, PQB_LOW}; const int o[] = {PQB_OUT_MIDPRICE};
    pqb_out_col *d[] = {out};
    // nulls in `low` make the reference fail (overlap.rs:352-376); nulls in `high` alone would
    // need a per-field start: both are refused.
    return run_single(e, c, f, 2, NP_ERR, &sp, o, d, 1);
}
