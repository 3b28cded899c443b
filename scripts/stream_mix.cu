// stream_mix.cu -- what HBM delivers for the suite's access mix, without the suite: a streaming kernel that reads 4 planes
// and writes 21 (the fused suite's 1 : 5.25 read : write ratio) with the same 256-byte warp rows, against a 1 : 1 copy.
// Measurement helper only (scripts/): nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/stream_mix scripts/stream_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct Planes { const double *in[4]; double *out[21]; };

template <int NIN, int NOUT, int VEC>
__global__ void __launch_bounds__(256) mix_kernel(const __grid_constant__ Planes P, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < n; i += stride) {
        double s[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] = 0.0;
#pragma unroll
        for (int f = 0; f < NIN; ++f) {
            if (VEC == 2) { const double2 x = __ldcs(reinterpret_cast<const double2 *>(P.in[f] + i)); s[0] += x.x; s[VEC - 1] += x.y; }
            else s[0] += __ldcs(P.in[f] + i);
        }
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
            if (VEC == 2) __stcs(reinterpret_cast<double2 *>(P.out[k] + i), make_double2(s[0] + k, s[VEC - 1] + k));
            else __stcs(P.out[k] + i, s[0] + k);
        }
    }
}

template <int NIN, int NOUT, int VEC>
static int run(const char *tag, const Planes &P, size_t n, int grid) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) mix_kernel<NIN, NOUT, VEC><<<grid, 256>>>(P, n);
    CK(cudaEventRecord(a));
    const int iters = 5;
    for (int i = 0; i < iters; ++i) mix_kernel<NIN, NOUT, VEC><<<grid, 256>>>(P, n);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    ms /= iters;
    printf("{\"kernel\": \"%s\", \"reads\": %d, \"writes\": %d, \"vector_bytes\": %d, \"grid\": %d, \"ms\": %.4f, \"gbs\": %.1f}\n", tag, NIN, NOUT,
           VEC * 8, grid, ms, (double)(NIN + NOUT) * n * 8 / ms / 1e6);
    return 0;
}

int main(int argc, char **argv) {
    const size_t n = (argc > 1 ? atoll(argv[1]) : 20000ll) * 5040;      // doubles per plane
    Planes P{};
    for (auto &p : P.in) { double *q; CK(cudaMalloc(&q, n * 8)); CK(cudaMemset(q, 0, n * 8)); p = q; }
    for (auto &p : P.out) CK(cudaMalloc(&p, n * 8));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (int per_sm : {4, 8}) {
        const int grid = sms * per_sm;
        if (run<4, 21, 1>("suite mix 4r/21w, 8-byte", P, n, grid)) return 1;
        if (run<4, 21, 2>("suite mix 4r/21w, 16-byte", P, n, grid)) return 1;
        if (run<1, 1, 2>("copy 1r/1w, 16-byte", P, n, grid)) return 1;
        if (run<4, 4, 2>("4r/4w, 16-byte", P, n, grid)) return 1;
        if (run<1, 5, 2>("1r/5w, 16-byte", P, n, grid)) return 1;
        if (run<0, 21, 2>("write only 21w, 16-byte", P, n, grid)) return 1;
        if (run<4, 1, 2>("4r/1w, 16-byte", P, n, grid)) return 1;
    }
    return 0;
}
