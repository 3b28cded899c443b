// probe: achievable HBM bandwidth of the "one thread per symbol row" access pattern
// (each thread streams its own row: 32 B loads of NIN planes, 32 B stores of NOUT planes per step)
// versus a coalesced warp-per-row pattern.  Not product code.
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

// thread per row; WS warps of a CTA share the same 32 rows and split the NOUT outputs among them
template<int WS>
__global__ void __launch_bounds__(32*WS) rowthread(const __grid_constant__ Args A){
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  const int s=blockIdx.x*32+lane;
  if(s>=A.S) return;
  const size_t row=(size_t)s*A.pitch;
  const int o0=(NOUT*w)/WS, o1=(NOUT*(w+1))/WS;
  double x[NIN][4], y[NIN][4];
  #pragma unroll
  for(int f=0;f<NIN;++f) ldv4(A.in[f]+row,x[f][0],x[f][1],x[f][2],x[f][3]);
  double acc=0;
  for(int t=0;t<A.N;t+=4){
    const int tn=(t+4<A.N)?t+4:t;
    #pragma unroll
    for(int f=0;f<NIN;++f) ldv4(A.in[f]+row+tn,y[f][0],y[f][1],y[f][2],y[f][3]);
    double v0=x[0][0]+x[1][0]+x[2][0]+x[3][0]+acc, v1=x[0][1]+x[1][1]+x[2][1]+x[3][1]+v0;
    double v2=x[0][2]+x[1][2]+x[2][2]+x[3][2]+v1, v3=x[0][3]+x[1][3]+x[2][3]+x[3][3]+v2; acc=v3*0.5;
    for(int k=o0;k<o1;++k) stv4(A.out[k]+row+t,v0+k,v1,v2,v3);
    #pragma unroll
    for(int f=0;f<NIN;++f){x[f][0]=y[f][0];x[f][1]=y[f][1];x[f][2]=y[f][2];x[f][3]=y[f][3];}
  }
}
// coalesced: warp per row, lane holds 4 consecutive bars (1 KB per plane per step)
__global__ void __launch_bounds__(256) rowwarp(const __grid_constant__ Args A){
  const int lane=threadIdx.x&31; const int gw=(blockIdx.x*blockDim.x+threadIdx.x)>>5; const int nw=(gridDim.x*blockDim.x)>>5;
  for(int s=gw;s<A.S;s+=nw){
    const size_t row=(size_t)s*A.pitch; double acc=0;
    for(int t=4*lane;t<A.N;t+=128){
      double x[NIN][4];
      #pragma unroll
      for(int f=0;f<NIN;++f) ldv4(A.in[f]+row+t,x[f][0],x[f][1],x[f][2],x[f][3]);
      double v0=x[0][0]+x[1][0]+x[2][0]+x[3][0]+acc, v1=x[0][1]+x[1][1]+x[2][1]+x[3][1]+v0;
      double v2=x[0][2]+x[1][2]+x[2][2]+x[3][2]+v1, v3=x[0][3]+x[1][3]+x[2][3]+x[3][3]+v2; acc=v3*0.5;
      #pragma unroll
      for(int k=0;k<NOUT;++k) stv4(A.out[k]+row+t,v0+k,v1,v2,v3);
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
  auto run=[&](const char*name,auto launch){
    for(int i=0;i<3;++i) launch(); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); for(int i=0;i<10;++i) launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=10;
    printf("%-28s S=%d N=%d  %.3f ms  %.1f GB/s\n",name,S,N,ms,bytes/ms/1e6);
  };
  int g=(S+31)/32;
  run("rowthread WS=1",[&]{rowthread<1><<<g,32>>>(A);});
  run("rowthread WS=3",[&]{rowthread<3><<<g,96>>>(A);});
  run("rowthread WS=7",[&]{rowthread<7><<<g,224>>>(A);});
  run("rowwarp grid=148*8",[&]{rowwarp<<<148*8,256>>>(A);});
  run("rowwarp grid=148*4",[&]{rowwarp<<<148*4,256>>>(A);});
  CK(cudaGetLastError());
  return 0;
}
