// probe/out_layout.cu -- does the LAYOUT of the 21 output planes limit what HBM delivers to the fused suite's access pattern?
// One CTA per symbol block (32 symbols), 7 "role" warps; per bar a block reads 4 x 256 B (tiled input planes, one warp each for
// the first four) and every role warp writes 3 x 256 B.  Outputs either PLANAR (21 planes, element (b, t, lane) of plane k at
// ((b * T + t) * 32 + lane), the engine's layout) or INTERLEAVED per block ([b][t][k][32]: one sequential stream per block).
// A short dependent FMA chain per bar stands in for the arithmetic.  nvcc -O3 -arch=sm_100a probe/out_layout.cu -o probe/out_layout
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr int NO = 21, NI = 4, NW = 7;
template <bool INTER, int CHAIN>
__global__ void __launch_bounds__(32 * NW, 3) k(const double *in, double *out, int T, size_t plane) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = blockIdx.x;
    const double *ip = in + (size_t)(w % NI) * plane + (size_t)b * T * 32 + lane;
    double acc = 0.0;
    for (int t = 0; t < T; ++t) {
        double x = __ldcs(ip + (size_t)t * 32);
#pragma unroll
        for (int c = 0; c < CHAIN; ++c) acc = fma(acc, 0.999, x);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int kk = w * 3 + j;
            double *op = INTER ? out + (((size_t)b * T + t) * NO + kk) * 32 + lane : out + (size_t)kk * plane + ((size_t)b * T + t) * 32 + lane;
            __stcs(op, acc + j);
        }
    }
}
template <bool INTER, int CHAIN>
float run(const double *in, double *out, int B, int T, size_t plane, int iters) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k<INTER, CHAIN><<<B, 32 * NW>>>(in, out, T, plane);
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) k<INTER, CHAIN><<<B, 32 * NW>>>(in, out, T, plane);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}
int main(int argc, char **argv) {
    const int T = 5040;
    for (int B : {444, 1563}) {
        const size_t plane = (size_t)B * T * 32;
        double *in, *out;
        CK(cudaMalloc(&in, plane * NI * 8)); CK(cudaMalloc(&out, plane * NO * 8));
        CK(cudaMemset(in, 0, plane * NI * 8));
        const double gb = (double)plane * (NI + NO) * 8 / 1e9;
        float p8 = run<false, 8>(in, out, B, T, plane, 5), i8 = run<true, 8>(in, out, B, T, plane, 5);
        float p24 = run<false, 24>(in, out, B, T, plane, 5), i24 = run<true, 24>(in, out, B, T, plane, 5);
        printf("blocks %4d: chain 8  planar %.3f ms %.0f GB/s | interleaved %.3f ms %.0f GB/s\n", B, p8, gb / p8 * 1e3, i8, gb / i8 * 1e3);
        printf("blocks %4d: chain 24 planar %.3f ms %.0f GB/s | interleaved %.3f ms %.0f GB/s\n", B, p24, gb / p24 * 1e3, i24, gb / i24 * 1e3);
        cudaFree(in); cudaFree(out);
    }
    return 0;
}
