// probe: store cache hints for the 4-read / 21-write streaming mix on the tiled layout
// (warp writes 256 contiguous bytes per instruction, like suite_fused_kernel).  Not product code.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)
constexpr int NIN=4, NOUT=21;
struct Args { const double* in[NIN]; double* out[NOUT]; int blocks, bars; };
template<int H> __device__ __forceinline__ void st(double*p,double v){
  if(H==0) asm volatile("st.global.f64 [%0], %1;"::"l"(p),"d"(v):"memory");
  if(H==1) asm volatile("st.global.cs.f64 [%0], %1;"::"l"(p),"d"(v):"memory");
  if(H==2) asm volatile("st.global.cg.f64 [%0], %1;"::"l"(p),"d"(v):"memory");
  if(H==3) asm volatile("st.global.wt.f64 [%0], %1;"::"l"(p),"d"(v):"memory");
  if(H==4) asm volatile("st.global.L1::no_allocate.f64 [%0], %1;"::"l"(p),"d"(v):"memory");
}
// one warp per (block, output subset): 7 warps split the 21 outputs, lane = symbol, serial over bars
template<int H>
__global__ void __launch_bounds__(224) k(const __grid_constant__ Args A){
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  const size_t base=(size_t)blockIdx.x*A.bars*32+lane;
  double acc=0;
  for(int t=0;t<A.bars;++t){
    const size_t o=base+(size_t)t*32;
    double x=A.in[0][o]+A.in[1][o]+A.in[2][o]+A.in[3][o];
    acc=acc*0.5+x;
    #pragma unroll
    for(int j=0;j<3;++j) st<H>(A.out[w*3+j]+o,acc+j);
  }
}
int main(int argc,char**argv){
  int S=argc>1?atoi(argv[1]):50000, N=argc>2?atoi(argv[2]):5040; int blocks=(S+31)/32;
  Args A; A.blocks=blocks; A.bars=N; size_t plane=(size_t)blocks*N*32*8;
  for(int f=0;f<NIN;++f){ double*p; CK(cudaMalloc(&p,plane)); CK(cudaMemset(p,0,plane)); A.in[f]=p; }
  for(int q=0;q<NOUT;++q){ CK(cudaMalloc(&A.out[q],plane)); }
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes=(double)blocks*32*N*8*(NIN+NOUT);
  auto run=[&](const char*name,auto launch){
    for(int i=0;i<2;++i) launch(); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); for(int i=0;i<5;++i) launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=5;
    printf("%-22s %.3f ms  %.1f GB/s\n",name,ms,bytes/ms/1e6);
  };
  run("st default",[&]{k<0><<<blocks,224>>>(A);});
  run("st.cs",[&]{k<1><<<blocks,224>>>(A);});
  run("st.cg",[&]{k<2><<<blocks,224>>>(A);});
  run("st.wt",[&]{k<3><<<blocks,224>>>(A);});
  run("st L1::no_allocate",[&]{k<4><<<blocks,224>>>(A);});
  CK(cudaGetLastError());
  return 0;
}
