// dev probe: DFMA throughput on B200 and whether it overlaps with IMAD.WIDE (can FP64 carry part of a bigint product?)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_IT 2048
__global__ void k_dfma(double* d) { double a[8]; for (int i=0;i<8;i++) a[i]=d[threadIdx.x]+i; double m=a[0]*1.0000001, c=a[1];
  for (int k=0;k<N_IT;k++) { _Pragma("unroll") for (int i=0;i<8;i++) a[i]=fma(a[i],m,c); } double r=0; for (int i=0;i<8;i++) r+=a[i]; d[blockIdx.x*blockDim.x+threadIdx.x]=r; }
__global__ void k_wide(uint32_t* d) { uint64_t a[8]; for (int i=0;i<8;i++) a[i]=d[threadIdx.x]+i; uint32_t m=(uint32_t)a[0]|1;
  for (int k=0;k<N_IT;k++) { _Pragma("unroll") for (int i=0;i<8;i++) a[i]=(uint64_t)(uint32_t)(a[i]>>7)*m+a[i]; } uint64_t r=0; for (int i=0;i<8;i++) r^=a[i]; d[blockIdx.x*blockDim.x+threadIdx.x]=(uint32_t)(r^(r>>32)); }
__global__ void k_both(uint32_t* d, double* e) { uint64_t a[4]; double f[8]; for (int i=0;i<4;i++) a[i]=d[threadIdx.x]+i; for (int i=0;i<8;i++) f[i]=e[threadIdx.x]+i;
  uint32_t m=(uint32_t)a[0]|1; double fm=f[0]*1.0000001, fc=f[1];
  for (int k=0;k<N_IT;k++) { _Pragma("unroll") for (int i=0;i<4;i++) { a[i]=(uint64_t)(uint32_t)(a[i]>>7)*m+a[i]; f[2*i]=fma(f[2*i],fm,fc); f[2*i+1]=fma(f[2*i+1],fm,fc); } }
  uint64_t r=0; double s=0; for (int i=0;i<4;i++) r^=a[i]; for (int i=0;i<8;i++) s+=f[i]; d[blockIdx.x*blockDim.x+threadIdx.x]=(uint32_t)(r^(r>>32)); e[blockIdx.x*blockDim.x+threadIdx.x]=s; }
template <class F> float timeit(F f) { cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); return ms; }
int main() { int blocks=148*8, threads=256; size_t n=(size_t)blocks*threads; uint32_t* d; double* e; cudaMalloc(&d,n*4); cudaMalloc(&e,n*8); cudaMemset(d,0x11,n*4); cudaMemset(e,0x3f,n*8);
  double sc = 1.0/148/1.965e9*1e3;
  float t1=timeit([&]{k_dfma<<<blocks,threads>>>(e);}); printf("dfma      %7.3f ms  %6.2f dfma/clk/SM\n", t1, (double)n*N_IT*8/t1*sc);
  float t2=timeit([&]{k_wide<<<blocks,threads>>>(d);}); printf("imad.wide %7.3f ms  %6.2f wide/clk/SM\n", t2, (double)n*N_IT*8/t2*sc);
  float t3=timeit([&]{k_both<<<blocks,threads>>>(d,e);}); printf("both      %7.3f ms  %6.2f wide + %6.2f dfma /clk/SM\n", t3, (double)n*N_IT*4/t3*sc, (double)n*N_IT*8/t3*sc);
  return 0; }
