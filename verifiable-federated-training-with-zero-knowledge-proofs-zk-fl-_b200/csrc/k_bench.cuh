// integer-pipe microbenchmark kernels (roofline denominators); only zkfl.cu includes it.
#pragma once
#include "types.cuh"

namespace zk {

// integer-pipe microbenchmark: `iters` dependent Montgomery products per thread (roofline denominator)
ZK_GLOBAL void k_bench_modmul(Fq* __restrict__ data, size_t n, uint32_t iters) {
  size_t i = ZK_TID;
  if (i >= n) return;
  Fq x = data[i], y = x;
  ZK_NOUNROLL for (uint32_t k = 0; k < iters; k++) { x = x * y; y = y * x; }
  data[i] = x + y;
}

// 32x32->64 multiply-accumulate microbenchmark: 4 (mad.lo.cc, madc.hi.cc) pairs per step on adjacent words, the form
// the Montgomery product is made of (ptxas fuses each pair into one IMAD.WIDE.U32.X). 4 wide MACs per thread per step.
ZK_GLOBAL void k_bench_widemac(uint32_t* __restrict__ data, size_t n, uint32_t iters) {
  size_t i = ZK_TID;
  if (i >= n) return;
  uint32_t a0 = data[i], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const uint32_t m = a0 | 1u, x = a1 | 3u, y = m ^ x;
  ZK_NOUNROLL for (uint32_t k = 0; k < iters; k += 4) {   // 4 steps per trip: keeps loop overhead out of the rate
#ifdef ZKFL_PTX_MUL
#define ZK_WIDEMAC_STEP                                                                                                          \
    asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;" \
                 "madc.lo.cc.u32 %4, %9, %10, %4; madc.hi.cc.u32 %5, %9, %10, %5; madc.lo.cc.u32 %6, %8, %8, %6; madc.hi.u32 %7, %8, %8, %7;"   \
                 : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m), "r"(x), "r"(y));
    ZK_WIDEMAC_STEP ZK_WIDEMAC_STEP ZK_WIDEMAC_STEP ZK_WIDEMAC_STEP
#undef ZK_WIDEMAC_STEP
#else
    for (int q = 0; q < 4; q++) {
      uint64_t t = (uint64_t)m * x + a0; a0 = (uint32_t)t; a1 += (uint32_t)(t >> 32);
      t = (uint64_t)m * y + a2; a2 = (uint32_t)t; a3 += (uint32_t)(t >> 32);
      t = (uint64_t)x * y + a4; a4 = (uint32_t)t; a5 += (uint32_t)(t >> 32);
      t = (uint64_t)m * m + a6; a6 = (uint32_t)t; a7 += (uint32_t)(t >> 32);
    }
#endif
  }
  data[i] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
// raw IMAD-chain microbenchmark: 8 independent 32-bit multiply-add chains per thread, iters steps each
ZK_GLOBAL void k_bench_imad(uint32_t* __restrict__ data, size_t n, uint32_t iters) {
  size_t i = ZK_TID;
  if (i >= n) return;
  uint32_t a0 = data[i], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const uint32_t m = a0 | 1u, c = a0 ^ 0x9e3779b9u;
  ZK_NOUNROLL for (uint32_t k = 0; k < iters; k++) {
    a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
    a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
  }
  data[i] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}


}  // namespace zk
