// probe: HBM bandwidth vs contiguous chunk size per row when a warp owns 32 rows and walks time
// in steps of T bars (the global access pattern of a thread-per-symbol kernel with smem transposes).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)
constexpr int NIN=4, NOUT=21;
struct Args { const double* in[NIN]; double* out[NOUT]; int S, N, pitch; };
__device__ __forceinline__ void ldv4(const double*p,double&a,double&b,double&c,double&d){
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];":"=d"(a),"=d"(b),"=d"(c),"=d"(d):"l"(p));}
__device__ __forceinline__ void stv4(double*p,double a,double b,double c,double d){
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"::"l"(p),"d"(a),"d"(b),"d"(c),"d"(d):"memory");}
// T bars per row per step; a warp instruction covers RPI = 128/T rows x T bars
template<int T>
__global__ void __launch_bounds__(256) chunked(const __grid_constant__ Args A, int rows_per_warp){
  constexpr int LPR=T/4;            // lanes per row
  constexpr int RPI=32/LPR;         // rows per instruction
  const int lane=threadIdx.x&31; const int gw=(blockIdx.x*blockDim.x+threadIdx.x)>>5; const int nw=(gridDim.x*blockDim.x)>>5;
  const int ngroups=(A.S+rows_per_warp-1)/rows_per_warp;
  const int rsub=lane/LPR, col=(lane%LPR)*4;
  for(int g=gw; g<ngroups; g+=nw){
    const int s0=g*rows_per_warp;
    for(int t0=0;t0<A.N;t0+=T){
      for(int r=rsub;r<rows_per_warp;r+=RPI){
        const int s=s0+r; if(s>=A.S||t0+col>=A.pitch) continue;
        const size_t off=(size_t)s*A.pitch+t0+col;
        double x[NIN][4];
        #pragma unroll
        for(int f=0;f<NIN;++f) ldv4(A.in[f]+off,x[f][0],x[f][1],x[f][2],x[f][3]);
        double v0=x[0][0]+x[1][0]+x[2][0]+x[3][0], v1=x[0][1]+x[1][1]+x[2][1]+x[3][1];
        double v2=x[0][2]+x[1][2]+x[2][2]+x[3][2], v3=x[0][3]+x[1][3]+x[2][3]+x[3][3];
        #pragma unroll
        for(int k=0;k<NOUT;++k) stv4(A.out[k]+off,v0+k,v1,v2,v3);
      }
    }
  }
}
int main(int argc,char**argv){
  int S=argc>1?atoi(argv[1]):50000, N=argc>2?atoi(argv[2]):5040; int pitch=(N+15)/16*16;
  Args A; A.S=S;A.N=N;A.pitch=pitch; size_t plane=(size_t)S*pitch*8;
  for(int f=0;f<NIN;++f){ double*p; CK(cudaMalloc(&p,plane)); CK(cudaMemset(p,0,plane)); A.in[f]=p; }
  for(int k=0;k<NOUT;++k){ CK(cudaMalloc(&A.out[k],plane)); }
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes=(double)S*N*8*(NIN+NOUT);
  auto run=[&](const char*name,int rpw,auto launch){
    for(int i=0;i<3;++i) launch(); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); for(int i=0;i<10;++i) launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=10;
    printf("%-20s rows/warp=%d S=%d N=%d  %.3f ms  %.1f GB/s\n",name,rpw,S,N,ms,bytes/ms/1e6);
  };
  for(int grid: {148*4,148*8}){
    printf("grid=%d\n",grid);
    run("T=16 (128B)",32,[&]{chunked<16><<<grid,256>>>(A,32);});
    run("T=32 (256B)",32,[&]{chunked<32><<<grid,256>>>(A,32);});
    run("T=64 (512B)",32,[&]{chunked<64><<<grid,256>>>(A,32);});
    run("T=128 (1KB)",32,[&]{chunked<128><<<grid,256>>>(A,32);});
    run("T=128 (1KB)",1,[&]{chunked<128><<<grid,256>>>(A,1);});
  }
  CK(cudaGetLastError());
  return 0;
}
