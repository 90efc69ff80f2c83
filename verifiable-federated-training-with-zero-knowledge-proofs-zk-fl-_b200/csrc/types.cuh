// Plain structs shared by the kernels and the host side of every translation unit of libzkfl.so.
#pragma once
#include "bn254.cuh"

namespace zk {


#ifndef ZKFL_EMUL
#define ZK_ATOMIC_MIN(p, v) atomicMin((p), (v))
#else
static inline void zk_atomic_min_u32(uint32_t* p, uint32_t v) {
  uint32_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
}
#define ZK_ATOMIC_MIN(p, v) zk_atomic_min_u32((p), (v))
#endif

// software prefetch of a line that a later iteration gathers (no register cost; a no-op in the host emulation)
#if defined(__CUDA_ARCH__)
#define ZK_PREFETCH(ptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr))
#else
#define ZK_PREFETCH(ptr) ((void)(ptr))
#endif

// occupancy of the bucket accumulation: 128 threads per CTA; G1 fits 4 CTAs per SM at 128 registers; the G2 kernel compiles to
// 238 registers (2 CTAs, 1.9 warps per scheduler in ncu) unless bounded (ZKFL_G2_MIN_CTAS = 3 caps it at 168)
#ifndef ZKFL_G2_MIN_CTAS
#define ZKFL_G2_MIN_CTAS 3   // measured: 3 -> 78.5 ms, unbounded -> 80.3 ms, 4 (128 registers, spills) -> 98.8 ms per 1024 proofs
#endif
#if defined(__CUDACC__) && !defined(ZKFL_EMUL)
#define ZK_ACC_BOUNDS(F) __launch_bounds__(128, sizeof(F) > 32 ? ZKFL_G2_MIN_CTAS : 4)
#define ZK_FIX_BOUNDS(F, BOUND) __launch_bounds__(128, (BOUND) == 1 ? (sizeof(F) > 32 ? 3 : 4) : 1)   // BOUND 2: few rows, inlined products, no cap
#define ZK_LVL_BOUNDS(F, BOUND) __launch_bounds__(64, (BOUND) ? (sizeof(F) > 32 ? 6 : 8) : 1)
#else
#define ZK_ACC_BOUNDS(F)
#define ZK_FIX_BOUNDS(F, BOUND)
#define ZK_LVL_BOUNDS(F, BOUND)
#endif

struct PoseidonDev {
  uint32_t rounds, rp;
  const Fr* C;  // rounds*t round constants, Montgomery
  const Fr* M;  // t*t MDS, row-major, Montgomery
};

struct ProgramDev {
  uint32_t n_wires, n_inputs, n_ops;
  const uint32_t* ops;       // 5 words per op
  const uint32_t* lc_off;
  const uint32_t* lc_wire;
  const Fr* lc_coef;         // coef * R^2: coef (*) canonical witness = Montgomery(coef * w)
  const uint32_t* pos_in;
  PoseidonDev pk[18];
};

struct CsrDev {
  const uint32_t* row_off;  // n_rows + 1
  const uint32_t* wire;
  const Fr* coef;           // coef * R^2 (the bytes zkey section 4 stores)
};

struct VkDev {
  G1Affine alpha1, beta1, delta1;
  G2Affine beta2, delta2;
};
typedef uint32_t zk_key_t;   // bucket index of a sorted entry (32 bits: the standalone MSM over resident bases uses up to 2^19 buckets)

// Batched over B proofs that share the bases. Signed c-bit digits: W = 254/c + 1 windows,
// nb = 2^(c-1) buckets per bucket set, row = b*R + (R == 1 ? 0 : j) identifies one bucket set.
struct MsmShape {
  uint32_t m;    // points
  uint32_t B;    // proofs
  uint32_t c, W, nb;
  uint32_t R;    // bucket sets ("rows") per proof: W (one per window) or 1 (all windows share one set: the bases
                 // table then holds 2^(c*j) * P_i at index j*m + i, so no doublings are needed after the reduction)
  uint32_t cap;  // entries reserved per row in the sorted index list (m, or m*W when R == 1)
  uint32_t lsS;  // 0: the sorted list of a row is linear. Otherwise it is stored CHUNK-TRANSPOSED for the batch-affine
                 // accumulation (chunks of S = 1 << lsS entries, cap a multiple of 32*S): entry r of chunk c sits at
                 // (c/32)*32*S + r*32 + c%32, so the 32 lanes of a warp (32 consecutive chunks) read consecutive words.
};

}  // namespace zk
