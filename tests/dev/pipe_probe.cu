// dev probe: integer-pipe throughput of the instruction forms a Montgomery product can be built from (sm_100a)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_IT 2048
__global__ void k_lo(uint32_t* d) { uint32_t a[8]; for (int i=0;i<8;i++) a[i]=d[threadIdx.x]+i; uint32_t m=a[0]|1, c=a[1];
  for (int k=0;k<N_IT;k++) { _Pragma("unroll") for (int i=0;i<8;i++) a[i]=a[i]*m+c; } uint32_t r=0; for (int i=0;i<8;i++) r^=a[i]; d[blockIdx.x*blockDim.x+threadIdx.x]=r; }
__global__ void k_hi(uint32_t* d) { uint32_t a[8]; for (int i=0;i<8;i++) a[i]=d[threadIdx.x]+i; uint32_t m=a[0]|0x80000001u, c=a[1];
  for (int k=0;k<N_IT;k++) { _Pragma("unroll") for (int i=0;i<8;i++) a[i]=__umulhi(a[i],m)+c; } uint32_t r=0; for (int i=0;i<8;i++) r^=a[i]; d[blockIdx.x*blockDim.x+threadIdx.x]=r; }
__global__ void k_wide(uint32_t* d) { uint64_t a[8]; for (int i=0;i<8;i++) a[i]=d[threadIdx.x]+i; uint32_t m=(uint32_t)a[0]|1;
  for (int k=0;k<N_IT;k++) { _Pragma("unroll") for (int i=0;i<8;i++) a[i]=(uint64_t)(uint32_t)a[i]*m+a[i]; } uint64_t r=0; for (int i=0;i<8;i++) r^=a[i]; d[blockIdx.x*blockDim.x+threadIdx.x]=(uint32_t)(r^(r>>32)); }
__global__ void k_add3(uint32_t* d) { uint32_t a[8]; for (int i=0;i<8;i++) a[i]=d[threadIdx.x]+i; uint32_t m=a[0]|1, c=a[1];
  for (int k=0;k<N_IT;k++) { _Pragma("unroll") for (int i=0;i<8;i++) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(a[i]) : "r"(m), "r"(c)); } uint32_t r=0; for (int i=0;i<8;i++) r^=a[i]; d[blockIdx.x*blockDim.x+threadIdx.x]=r; }
// lo multiply on the FMA pipe + independent adds on the ALU pipe: do the pipes overlap?
__global__ void k_mix(uint32_t* d) { uint32_t a[8], b[8]; for (int i=0;i<8;i++) { a[i]=d[threadIdx.x]+i; b[i]=a[i]*3; } uint32_t m=a[0]|1, c=a[1];
  for (int k=0;k<N_IT;k++) { _Pragma("unroll") for (int i=0;i<8;i++) { a[i]=a[i]*m+c; asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(c)); } } uint32_t r=0; for (int i=0;i<8;i++) r^=a[i]^b[i]; d[blockIdx.x*blockDim.x+threadIdx.x]=r; }
// carry chain: mad.lo.cc + madc.hi.cc pairs (what a 32-bit-limb row operation is made of)
__global__ void k_chain(uint32_t* d) { uint32_t a[8]; for (int i=0;i<8;i++) a[i]=d[threadIdx.x]+i; uint32_t m=a[0]|1, x=a[1]|3;
  for (int k=0;k<N_IT;k++) {
    asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                 "madc.lo.cc.u32 %4, %9, %10, %4; madc.hi.cc.u32 %5, %9, %10, %5; madc.lo.cc.u32 %6, %8, %8, %6; madc.hi.u32 %7, %9, %9, %7;"
                 : "+r"(a[0]),"+r"(a[1]),"+r"(a[2]),"+r"(a[3]),"+r"(a[4]),"+r"(a[5]),"+r"(a[6]),"+r"(a[7]) : "r"(m),"r"(x),"r"(m^x)); }
  uint32_t r=0; for (int i=0;i<8;i++) r^=a[i]; d[blockIdx.x*blockDim.x+threadIdx.x]=r; }
template <class K> void run(const char* name, K k, uint32_t* d, double ops_per_thread_iter) {
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks=148*8, threads=256;
  k<<<blocks,threads>>>(d); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<<<blocks,threads>>>(d); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double ops=(double)blocks*threads*N_IT*ops_per_thread_iter;
  printf("%-10s %8.3f ms  %8.2f Gop/s  = %6.2f thread-ops/clk/SM @1.965GHz\n", name, ms, ops/ms/1e6, ops/(ms*1e-3)/148/1.965e9);
}
int main() { uint32_t* d; cudaMalloc(&d, 148*8*256*4); cudaMemset(d, 0x11, 148*8*256*4);
  run("imad.lo", k_lo, d, 8); run("imad.hi", k_hi, d, 8); run("imad.wide", k_wide, d, 8); run("add x2", k_add3, d, 16);
  run("lo+add", k_mix, d, 16); run("cc-chain", k_chain, d, 8); return 0; }
