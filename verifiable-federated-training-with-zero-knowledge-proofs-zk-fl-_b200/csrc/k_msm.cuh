// sm_100a kernels of the Pippenger multi-scalar multiplications, templated over the coordinate field,
// plus the per-key precomputations and the small group-law kernels of setup / split proofs.  Template kernels are instantiated by
// msm_g1.cu (Fq) and msm_g2.cu (Fq2); the non-template sort kernels (section ZK_K_MSM_SORT) are compiled by msm_g1.cu only.
#pragma once
#include "types.cuh"

namespace zk {

// ================================================================================ K6/K7: Pippenger MSM
// Batched over B proofs that share the bases. Signed c-bit digits: W = 254/c + 1 windows,
// nb = 2^(c-1) buckets per bucket set, row = b*R + (R == 1 ? 0 : j) identifies one bucket set.
// address of sorted position `pos` inside a row's list region
ZK_HD uint32_t msm_list_index(const MsmShape& s, uint32_t pos) {
  if (s.lsS == 0) return pos;
  const uint32_t S = 1u << s.lsS, chunk = pos >> s.lsS, r = pos & (S - 1);
  return ((chunk >> 5) << (5 + s.lsS)) + (r << 5) + (chunk & 31u);
}
ZK_HD uint32_t scalar_bits(const uint32_t* k, uint32_t pos, uint32_t c) {
  uint32_t word = pos >> 5, off = pos & 31;
  if (word >= 8) return 0;
  uint32_t v = k[word] >> off;
  if (off + c > 32 && word < 7) v |= k[word + 1] << (32 - off);
  return v & ((1u << c) - 1);
}
// digit j of the signed recoding; carry is threaded by the caller across j = 0..W-1
ZK_HD int32_t signed_digit(const uint32_t* k, uint32_t j, uint32_t c, uint32_t& carry) {
  uint32_t d = scalar_bits(k, j * c, c) + carry;
  if (d > (1u << (c - 1))) { carry = 1; return (int32_t)d - (int32_t)(1u << c); }
  carry = 0;
  return (int32_t)d;
}

#ifdef ZK_K_MSM_SORT
// pass 1: bucket histogram. scalars: canonical [m][B]. counts: [B*W][nb]. skip[i] != 0 drops point i
// (bases that are the point at infinity: wires absent from the B matrix, public wires of the C query).
ZK_GLOBAL void k_msm_count(const Fr* __restrict__ scalars, const uint8_t* __restrict__ skip, MsmShape s,
                           uint32_t* __restrict__ counts) {
  // proof-major thread order: a CTA works inside ONE proof's counters / list region (about 1 MB), which stays in L2;
  // the scalar loads become 32-byte strided sectors (each scalar is exactly one sector, so no DRAM traffic is wasted).
  size_t tid = ZK_TID;
  if (tid >= (size_t)s.m * s.B) return;
  uint32_t b = (uint32_t)(tid / s.m), i = (uint32_t)(tid % s.m);
  if (skip && skip[i]) return;
  Fr k = scalars[(size_t)i * s.B + b];
  if (k.is_zero()) return;
  uint32_t carry = 0;
  for (uint32_t j = 0; j < s.W; j++) {
    int32_t d = signed_digit(k.v, j, s.c, carry);
    if (d == 0) continue;
    uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
    size_t row = s.R == 1 ? (size_t)b : (size_t)b * s.W + j;
    ZK_ATOMIC_ADD(counts + row * s.nb + (mag - 1), 1u);
  }
}
// pass 2a: per-row chunk sums (chunk = SCAN_CHUNK buckets)
#define ZK_SCAN_CHUNK 128u
ZK_GLOBAL void k_msm_scan_chunks(const uint32_t* __restrict__ counts, MsmShape s, uint32_t* __restrict__ chunk_sums) {
  size_t tid = ZK_TID;
  uint32_t nchunk = (s.nb + ZK_SCAN_CHUNK - 1) / ZK_SCAN_CHUNK;
  if (tid >= (size_t)s.B * s.R * nchunk) return;
  size_t row = tid / nchunk;
  uint32_t ch = (uint32_t)(tid % nchunk);
  uint32_t lo = ch * ZK_SCAN_CHUNK, hi = lo + ZK_SCAN_CHUNK < s.nb ? lo + ZK_SCAN_CHUNK : s.nb;
  uint32_t acc = 0;
  for (uint32_t k = lo; k < hi; k++) acc += counts[row * s.nb + k];
  chunk_sums[tid] = acc;
}
// pass 2b: exclusive offsets inside the row's region of the sorted list; cursors = copy used by the scatter
ZK_GLOBAL void k_msm_scan_write(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ chunk_sums, MsmShape s,
                                uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursors) {
  size_t tid = ZK_TID;
  uint32_t nchunk = (s.nb + ZK_SCAN_CHUNK - 1) / ZK_SCAN_CHUNK;
  if (tid >= (size_t)s.B * s.R * nchunk) return;
  size_t row = tid / nchunk;
  uint32_t ch = (uint32_t)(tid % nchunk);
  uint32_t acc = 0;
  for (uint32_t q = 0; q < ch; q++) acc += chunk_sums[row * nchunk + q];
  uint32_t lo = ch * ZK_SCAN_CHUNK, hi = lo + ZK_SCAN_CHUNK < s.nb ? lo + ZK_SCAN_CHUNK : s.nb;
  for (uint32_t k = lo; k < hi; k++) {
    offsets[row * s.nb + k] = acc;
    cursors[row * s.nb + k] = acc;
    acc += counts[row * s.nb + k];
  }
}
// pass 3: scatter point references into bucket order. sorted: [B*W][cap], entry = point | sign << 31;
// skey (may be NULL): the bucket index of every entry -- only the batch-affine accumulation reads it; the chunk kernel finds its
// runs from the offsets, which saves the second scattered store per entry (measured: the stores, not the atomics, bound this pass)
ZK_GLOBAL void k_msm_scatter(const Fr* __restrict__ scalars, const uint8_t* __restrict__ skip, MsmShape s,
                             uint32_t* __restrict__ cursors, uint32_t* __restrict__ sorted, zk_key_t* __restrict__ skey) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)s.m * s.B) return;
  uint32_t b = (uint32_t)(tid / s.m), i = (uint32_t)(tid % s.m);   // proof-major, see k_msm_count
  if (skip && skip[i]) return;
  Fr k = scalars[(size_t)i * s.B + b];
  if (k.is_zero()) return;
  uint32_t carry = 0;
  for (uint32_t j = 0; j < s.W; j++) {
    int32_t d = signed_digit(k.v, j, s.c, carry);
    if (d == 0) continue;
    uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
    size_t row = s.R == 1 ? (size_t)b : (size_t)b * s.W + j;
    uint32_t pos = ZK_ATOMIC_ADD(cursors + row * s.nb + (mag - 1), 1u);
    uint32_t ref = s.R == 1 ? j * s.m + i : i;
    const size_t at = (size_t)row * s.cap + msm_list_index(s, pos);
    sorted[at] = ref | (d < 0 ? 0x80000000u : 0u);
    if (skey) skey[at] = (zk_key_t)(mag - 1);
  }
}
#ifndef ZKFL_EMUL
// passes 1-3 in ONE kernel for the batch case (one bucket set per proof, many proofs): a CTA owns one proof, keeps its whole
// histogram / cursor array in SHARED memory (nb <= 32768 counters = 128 KB of the SM's 227 KB), and makes the two sweeps over
// the proof's scalars with shared-memory atomics instead of atomics on L2-resident counters: histogram, block-wide exclusive
// scan (offsets and counts go to global memory for the accumulation / reduction kernels), scatter.  The two-pass structure and the
// resulting lists are those of k_msm_count / k_msm_scan_* / k_msm_scatter (entries of one bucket in arbitrary order).
// grid = rows (R == 1: row = proof), block = 1024 threads, dynamic shared memory = (nb + 32) * 4 bytes.
static __global__ void __launch_bounds__(1024) k_msm_sort_cta(const Fr* __restrict__ scalars, const uint8_t* __restrict__ skip, MsmShape s,
                                                              uint32_t* __restrict__ offsets, uint32_t* __restrict__ counts,
                                                              uint32_t* __restrict__ sorted, zk_key_t* __restrict__ skey) {
  extern __shared__ uint32_t zk_sort_sm[];
  uint32_t* cnt = zk_sort_sm;            // nb counters, later the scatter cursors
  uint32_t* wsum = zk_sort_sm + s.nb;    // 32 warp totals of the scan
  const uint32_t b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31u, warp = tid >> 5;
  for (uint32_t k = tid; k < s.nb; k += nt) cnt[k] = 0;
  __syncthreads();
  for (uint32_t i = tid; i < s.m; i += nt) {
    if (skip && skip[i]) continue;
    const Fr k = scalars[(size_t)i * s.B + b];
    if (k.is_zero()) continue;
    uint32_t carry = 0;
    for (uint32_t j = 0; j < s.W; j++) {
      const int32_t d = signed_digit(k.v, j, s.c, carry);
      if (d == 0) continue;
      atomicAdd(cnt + ((d < 0 ? (uint32_t)(-d) : (uint32_t)d) - 1), 1u);
    }
  }
  __syncthreads();
  // exclusive scan of the nb counters: `per` consecutive counters per thread, warp shuffles, then the 32 warp totals
  const uint32_t per = (s.nb + nt - 1) / nt, lo = tid * per, hi = lo + per < s.nb ? lo + per : s.nb;
  uint32_t mine = 0;
  for (uint32_t k = lo; k < hi; k++) mine += cnt[k];
  uint32_t incl = mine;
  ZK_UNROLL for (uint32_t o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (nt >> 5) ? wsum[lane] : 0, wi = w;
    ZK_UNROLL for (uint32_t o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    wsum[lane] = wi - w;                 // exclusive prefix of the warp totals
  }
  __syncthreads();
  uint32_t run = wsum[warp] + incl - mine;
  uint32_t* off_g = offsets + (size_t)b * s.nb;
  uint32_t* cnt_g = counts + (size_t)b * s.nb;
  for (uint32_t k = lo; k < hi; k++) {
    const uint32_t n = cnt[k];
    off_g[k] = run; cnt_g[k] = n;
    cnt[k] = run;                        // becomes the cursor of bucket k
    run += n;
  }
  __syncthreads();
  uint32_t* list = sorted + (size_t)b * s.cap;
  zk_key_t* keys = skey + (size_t)b * s.cap;
  for (uint32_t i = tid; i < s.m; i += nt) {
    if (skip && skip[i]) continue;
    const Fr k = scalars[(size_t)i * s.B + b];
    if (k.is_zero()) continue;
    uint32_t carry = 0;
    for (uint32_t j = 0; j < s.W; j++) {
      const int32_t d = signed_digit(k.v, j, s.c, carry);
      if (d == 0) continue;
      const uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
      const uint32_t at = msm_list_index(s, atomicAdd(cnt + (mag - 1), 1u));
      list[at] = (j * s.m + i) | (d < 0 ? 0x80000000u : 0u);
      if (skey) keys[at] = (zk_key_t)(mag - 1);
    }
  }
}
#endif
#endif  // ZK_K_MSM_SORT
// pass 4: BALANCED bucket accumulation. One thread per (row, chunk of S consecutive sorted entries): every thread
// performs exactly S mixed adds whatever the bucket-size distribution (witness scalars are far from uniform:
// bits, small values, and the partial top window concentrate thousands of entries in a few buckets).
// A bucket lying inside one chunk is written directly; a bucket cut by chunk borders leaves partial sums
// (head = run containing the chunk's first entry, tail = run containing its last entry) for pass 4b.
template <class F>
ZK_GLOBAL ZK_ACC_BOUNDS(F) void k_msm_accumulate_chunks(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                       const zk_key_t* __restrict__ skey, const uint32_t* __restrict__ offsets,
                                       const uint32_t* __restrict__ counts, MsmShape s, uint32_t S, uint32_t chunks_per_row,
                                       Xyzz<F>* __restrict__ buckets, Xyzz<F>* __restrict__ head, Xyzz<F>* __restrict__ tail) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)s.B * s.R * chunks_per_row) return;
  size_t row = tid / chunks_per_row;
  uint32_t ch = (uint32_t)(tid % chunks_per_row);
  const uint32_t* off = offsets + row * s.nb;
  const uint32_t* cnt = counts + row * s.nb;
  uint32_t total = off[s.nb - 1] + cnt[s.nb - 1];
  uint32_t pos0 = ch * S;
  if (pos0 >= total) return;
  uint32_t pos1 = pos0 + S < total ? pos0 + S : total;
  const uint32_t* list = sorted + row * s.cap;
  // the run containing pos0: the last bucket whose offset is <= pos0 (offsets are exclusive prefix sums, so among buckets sharing
  // an offset the last one is the non-empty one); binary search over the row's nb offsets
  uint32_t lo = 0, hi = s.nb;                       // invariant: off[lo] <= pos0, (hi == nb or off[hi] > pos0)
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (off[mid] <= pos0) lo = mid; else hi = mid; }
  uint32_t cur = lo, run_end = off[cur] + cnt[cur];
  bool first = true;
  Xyzz<F> acc = Xyzz<F>::infinity();
  // one entry of lookahead: the next list word is loaded, and the next base prefetched, while this entry's mixed add runs
  uint32_t e_next = list[msm_list_index(s, pos0)], e_next2 = 0;
  if (pos0 + 1 < pos1) e_next2 = list[msm_list_index(s, pos0 + 1)];
  for (uint32_t pos = pos0; pos < pos1; pos++) {
    const uint32_t e = e_next;
    e_next = e_next2;
    if (pos + 1 < pos1) {
      ZK_PREFETCH(bases + (e_next & 0x7FFFFFFFu));
      if (pos + 2 < pos1) e_next2 = list[msm_list_index(s, pos + 2)];
    }
    if (pos == run_end) {
      // the run of bucket `cur` ends inside this chunk; it is whole unless it began in an earlier chunk
      if (first && off[cur] < pos0) head[tid] = acc; else buckets[row * s.nb + cur] = acc;
      acc = Xyzz<F>::infinity();
      do { cur++; } while (cnt[cur] == 0);           // pos < total: a later non-empty bucket exists
      run_end = off[cur] + cnt[cur];
      first = false;
    }
    xyzz_madd(acc, bases[e & 0x7FFFFFFFu], (e >> 31) != 0);
  }
  bool starts_here = !(first && off[cur] < pos0);
  bool ends_here = run_end <= pos1;
  if (starts_here && ends_here) buckets[row * s.nb + cur] = acc;
  else if (first) head[tid] = acc;   // run covers the chunk's first entry (possibly the whole chunk)
  else tail[tid] = acc;              // run started here and continues in the next chunk
}
// pass 4 for G2, OPERAND-FILE form (OPT-IN, ZKFL_G2_OPERAND_FILE=1; bit-exact in the tests).  ncu of k_msm_accumulate_chunks<Fq2> (profiles/r02_sass_opcode_mix_accumulate.txt):
// 7 % of its executed instructions are IMAD.MOV / IMAD.U32 marshalling the 48 words of every by-value Fq2 product call and 2 % are
// spills -- all on the FMA pipe that bounds the kernel.  Here the running sum and every temporary of the mixed addition live in a
// per-thread OPERAND FILE in shared memory (9 Fq2 slots, 128-bit accesses, [slot][quarter][thread]: conflict-free) and ONE generic
// operation  dst = (a [- pre]) * b [- p1] [- 2 p2]  (or the square of the first factor) is a call that takes three integers: nothing
// is marshalled, nothing is live across the call, the file traffic (~14 LDS/STS.128 per operation) runs on the load/store pipe.
// The whole kernel is ~1.5 k instructions (one product body, one square body) instead of 4.5 k.
// The host emulation keeps the file in a thread-private array and runs the same operation sequence.
// MEASURED on B200 (1024 proofs, chunks of 64): 72.0 ms against 70.7 ms for the by-value-call kernel.  ncu of this form
// (profiles/r02_sass_opcode_mix_g2_operand_file.txt): IMAD.MOV 6.5 -> 3.2 % and no spills, but the slot address arithmetic brings
// IMAD 3.0 -> 5.0 % and the operand decoding LOP3 4.8 -> 6.5 %, the merges of the optional operands keep moves alive, and the IMAD.X
// carries (6.3 %) are untouched: non-multiply work on the FMA pipe 19.9 -> 17.4 % of the instructions, not enough to pay for the
// longer dependent chain through shared memory.  Kept as the starting point for a form with constant strides and branch-free operands.
enum { G2F_X = 0, G2F_Y, G2F_ZZ, G2F_ZZZ, G2F_P, G2F_R, G2F_PP, G2F_PPP, G2F_Q, G2F_SLOTS, G2F_NONE = 15 };
#ifndef ZKFL_EMUL
typedef uint4* G2FilePtr;                      // shared memory + threadIdx.x; slot s, quarter q at [(s * 4 + q) * blockDim.x]
__device__ __forceinline__ Fq2 g2f_ld(G2FilePtr f, uint32_t s) {
  const uint4* p = f + (size_t)s * 4 * blockDim.x;
  const uint4 v0 = p[0], v1 = p[blockDim.x], v2 = p[2 * blockDim.x], v3 = p[3 * blockDim.x];
  Fq2 r;
  r.a.v[0] = v0.x; r.a.v[1] = v0.y; r.a.v[2] = v0.z; r.a.v[3] = v0.w; r.a.v[4] = v1.x; r.a.v[5] = v1.y; r.a.v[6] = v1.z; r.a.v[7] = v1.w;
  r.b.v[0] = v2.x; r.b.v[1] = v2.y; r.b.v[2] = v2.z; r.b.v[3] = v2.w; r.b.v[4] = v3.x; r.b.v[5] = v3.y; r.b.v[6] = v3.z; r.b.v[7] = v3.w;
  return r;
}
__device__ __forceinline__ void g2f_st(G2FilePtr f, uint32_t s, const Fq2& x) {
  uint4* p = f + (size_t)s * 4 * blockDim.x;
  p[0] = make_uint4(x.a.v[0], x.a.v[1], x.a.v[2], x.a.v[3]); p[blockDim.x] = make_uint4(x.a.v[4], x.a.v[5], x.a.v[6], x.a.v[7]);
  p[2 * blockDim.x] = make_uint4(x.b.v[0], x.b.v[1], x.b.v[2], x.b.v[3]); p[3 * blockDim.x] = make_uint4(x.b.v[4], x.b.v[5], x.b.v[6], x.b.v[7]);
}
#else
typedef Fq2* G2FilePtr;
static inline Fq2 g2f_ld(G2FilePtr f, uint32_t s) { return f[s]; }
static inline void g2f_st(G2FilePtr f, uint32_t s, const Fq2& x) { f[s] = x; }
#endif
// code: dst | a << 4 | b << 8 | pre << 12 | p1 << 16 | p2 << 20 | sqr << 24 | neg_b << 25; b == G2F_NONE: second factor = *gb (global)
#define G2F_OP(dst, a, b, pre, p1, p2, sqr, neg) \
  ((uint32_t)(dst) | (uint32_t)(a) << 4 | (uint32_t)(b) << 8 | (uint32_t)(pre) << 12 | (uint32_t)(p1) << 16 | (uint32_t)(p2) << 20 | (uint32_t)(sqr) << 24 | (uint32_t)(neg) << 25)
// returns 1 when the result is zero
#ifndef ZKFL_EMUL
static __device__ __noinline__
#else
static inline
#endif
uint32_t g2f_op(G2FilePtr f, uint32_t code, const Fq2* gb) {
  const uint32_t dst = code & 15u, a = (code >> 4) & 15u, b = (code >> 8) & 15u, pre = (code >> 12) & 15u, p1 = (code >> 16) & 15u,
                 p2 = (code >> 20) & 15u;
  Fq2 x = g2f_ld(f, a);
  if (pre != G2F_NONE) x = x - g2f_ld(f, pre);
  Fq2 r;
  if ((code >> 24) & 1u) {
    const Fq t = Fq::mul_inline(x.a, x.b);
    r.a = Fq::mul_inline(x.a + x.b, x.a - x.b); r.b = t.dbl();
  } else {
    Fq2 y = b == G2F_NONE ? *gb : g2f_ld(f, b);
    if ((code >> 25) & 1u) y = y.neg();
    const Fq aa = Fq::mul_inline(x.a, y.a), bb = Fq::mul_inline(x.b, y.b), ss = Fq::mul_inline(x.a + x.b, y.a + y.b);
    r.a = aa - bb; r.b = ss - aa - bb;
  }
  if (p1 != G2F_NONE) r = r - g2f_ld(f, p1);
  if (p2 != G2F_NONE) r = r - g2f_ld(f, p2).dbl();
  g2f_st(f, dst, r);
  return r.is_zero() ? 1u : 0u;
}
// acc (in the file) += q (affine, global), exactly xyzz_madd; `inf`: the running sum is the point at infinity (nothing valid in the file)
ZK_D void g2f_madd(G2FilePtr f, bool& inf, const G2Affine* __restrict__ qp, bool negate) {
  const G2Affine q = *qp;
  if (q.is_inf()) return;
  if (inf) {
    g2f_st(f, G2F_X, q.x); g2f_st(f, G2F_Y, negate ? q.y.neg() : q.y); g2f_st(f, G2F_ZZ, Fq2::one()); g2f_st(f, G2F_ZZZ, Fq2::one());
    inf = false;
    return;
  }
  const uint32_t zp = g2f_op(f, G2F_OP(G2F_P, G2F_ZZ, G2F_NONE, G2F_NONE, G2F_X, G2F_NONE, 0, 0), &qp->x);        // P = x2 ZZ1 - X1
  const uint32_t zr = g2f_op(f, G2F_OP(G2F_R, G2F_ZZZ, G2F_NONE, G2F_NONE, G2F_Y, G2F_NONE, 0, negate ? 1 : 0), &qp->y);   // R = y2 ZZZ1 - Y1
  if (zp) {                                                   // same x: doubling or cancellation (never on the hot path)
    if (zr) {
      G2Affine qq = q; if (negate) qq.y = qq.y.neg();
      const G2Xyzz d = xyzz_dbl_affine(qq);
      g2f_st(f, G2F_X, d.X); g2f_st(f, G2F_Y, d.Y); g2f_st(f, G2F_ZZ, d.ZZ); g2f_st(f, G2F_ZZZ, d.ZZZ);
    } else inf = true;
    return;
  }
  g2f_op(f, G2F_OP(G2F_PP, G2F_P, G2F_NONE, G2F_NONE, G2F_NONE, G2F_NONE, 1, 0), nullptr);                          // PP = P^2
  g2f_op(f, G2F_OP(G2F_PPP, G2F_P, G2F_PP, G2F_NONE, G2F_NONE, G2F_NONE, 0, 0), nullptr);                           // PPP = P PP
  g2f_op(f, G2F_OP(G2F_Q, G2F_X, G2F_PP, G2F_NONE, G2F_NONE, G2F_NONE, 0, 0), nullptr);                             // Q = X1 PP
  g2f_op(f, G2F_OP(G2F_X, G2F_R, G2F_NONE, G2F_NONE, G2F_PPP, G2F_Q, 1, 0), nullptr);                               // X3 = R^2 - PPP - 2 Q
  g2f_op(f, G2F_OP(G2F_Y, G2F_Y, G2F_PPP, G2F_NONE, G2F_NONE, G2F_NONE, 0, 0), nullptr);                            // Y1 PPP
  g2f_op(f, G2F_OP(G2F_Y, G2F_Q, G2F_R, G2F_X, G2F_Y, G2F_NONE, 0, 0), nullptr);                                    // Y3 = (Q - X3) R - Y1 PPP
  g2f_op(f, G2F_OP(G2F_ZZ, G2F_ZZ, G2F_PP, G2F_NONE, G2F_NONE, G2F_NONE, 0, 0), nullptr);                           // ZZ3 = ZZ1 PP
  g2f_op(f, G2F_OP(G2F_ZZZ, G2F_ZZZ, G2F_PPP, G2F_NONE, G2F_NONE, G2F_NONE, 0, 0), nullptr);                        // ZZZ3 = ZZZ1 PPP
}
ZK_D G2Xyzz g2f_read(G2FilePtr f, bool inf) {
  if (inf) return G2Xyzz::infinity();
  G2Xyzz r; r.X = g2f_ld(f, G2F_X); r.Y = g2f_ld(f, G2F_Y); r.ZZ = g2f_ld(f, G2F_ZZ); r.ZZZ = g2f_ld(f, G2F_ZZZ);
  return r;
}
// same decomposition, outputs and run logic as k_msm_accumulate_chunks; dynamic shared memory = G2F_SLOTS * 64 * blockDim.x bytes
#if defined(__CUDACC__) && !defined(ZKFL_EMUL)
static __global__ void __launch_bounds__(128, 3)
#else
static void
#endif
k_msm_accumulate_chunks_g2f(const G2Affine* __restrict__ bases, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                            const uint32_t* __restrict__ counts, MsmShape s, uint32_t S, uint32_t chunks_per_row,
                            G2Xyzz* __restrict__ buckets, G2Xyzz* __restrict__ head, G2Xyzz* __restrict__ tail) {
#ifndef ZKFL_EMUL
  extern __shared__ uint4 zk_g2f_sm[];
  G2FilePtr f = zk_g2f_sm + threadIdx.x;
#else
  Fq2 file[G2F_SLOTS];
  G2FilePtr f = file;
#endif
  size_t tid = ZK_TID;
  if (tid >= (size_t)s.B * s.R * chunks_per_row) return;
  size_t row = tid / chunks_per_row;
  uint32_t ch = (uint32_t)(tid % chunks_per_row);
  const uint32_t* off = offsets + row * s.nb;
  const uint32_t* cnt = counts + row * s.nb;
  uint32_t total = off[s.nb - 1] + cnt[s.nb - 1];
  uint32_t pos0 = ch * S;
  if (pos0 >= total) return;
  uint32_t pos1 = pos0 + S < total ? pos0 + S : total;
  const uint32_t* list = sorted + row * s.cap;
  uint32_t lo = 0, hi = s.nb;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (off[mid] <= pos0) lo = mid; else hi = mid; }
  uint32_t cur = lo, run_end = off[cur] + cnt[cur];
  bool first = true, inf = true;
  uint32_t e_next = list[pos0], e_next2 = 0;
  if (pos0 + 1 < pos1) e_next2 = list[pos0 + 1];
  for (uint32_t pos = pos0; pos < pos1; pos++) {
    const uint32_t e = e_next;
    e_next = e_next2;
    if (pos + 1 < pos1) {
      ZK_PREFETCH(bases + (e_next & 0x7FFFFFFFu));
      if (pos + 2 < pos1) e_next2 = list[pos + 2];
    }
    if (pos == run_end) {
      if (first && off[cur] < pos0) head[tid] = g2f_read(f, inf); else buckets[row * s.nb + cur] = g2f_read(f, inf);
      inf = true;
      do { cur++; } while (cnt[cur] == 0);
      run_end = off[cur] + cnt[cur];
      first = false;
    }
    g2f_madd(f, inf, bases + (e & 0x7FFFFFFFu), (e >> 31) != 0);
  }
  bool starts_here = !(first && off[cur] < pos0);
  bool ends_here = run_end <= pos1;
  if (starts_here && ends_here) buckets[row * s.nb + cur] = g2f_read(f, inf);
  else if (first) head[tid] = g2f_read(f, inf);
  else tail[tid] = g2f_read(f, inf);
}
// pass 4, BATCH-AFFINE variant (large batches).  Same decomposition into chunks of S sorted entries and the same outputs
// (whole buckets written directly, head/tail partials for pass 4b), but the running sums stay AFFINE, and the S additions
// of a chunk are interleaved with those of the K-1 other chunks of the same thread and of the 32*K chunks of the warp, so
// that ONE field inversion serves 32*K additions (Montgomery's trick, shared across the warp).
//   sweep r = 0..S-1 over the thread's K slots (direction alternates):
//     finish addition r of the slot: 1/d from the prefix product stored by the previous sweep, lambda = (y_P - y_acc)/d,
//       x3 = lambda^2 - x_acc - x_P, y3 = lambda (x_acc - x3) - y_acc;
//     prepare addition r+1: d' = x_P' - x3, exclusive prefix product of the d' to scratch;
//   between sweeps: one inversion of the warp's total product (coop_inverse: warp scans + binary Euclid, warp-uniform).
// 6 products per addition plus the shared inversion and 12 scan products per K additions, instead of the 10 of the XYZZ
// mixed add; the price is ~200 B of coalesced scratch traffic per addition (HBM streams, prefetched one slot ahead).
// Doublings / cancellations / infinities are handled exactly (denominator substituted, never zero).
// Geometry: a row's list holds cpr = cap/S chunks = cpr32 groups of 32 chunks (one per lane); cpr32 is a multiple of K and a
// warp owns K consecutive groups of ONE row; slot (group G, lane) of the scratch arrays is G*32 + lane.
template <class F> ZK_D F coop_inverse(const F& total, uint32_t lane) {
#ifdef ZKFL_EMUL
  (void)lane;
  return total.inv_gcd();   // the emulation runs lanes one after the other: same value, no sharing
#else
  F incl = total, suf = total;
  ZK_UNROLL for (uint32_t off = 1; off < 32; off <<= 1) {
    F t = warp_shfl<ZK_SHFL_UP>(incl, off), u = warp_shfl<ZK_SHFL_DOWN>(suf, off);
    F mi = incl * t, ms = suf * u;
    if (lane >= off) incl = mi;
    if (lane + off < 32) suf = ms;
  }
  F all_inv = warp_shfl<ZK_SHFL_IDX>(incl, 31).inv_gcd();   // same value in every lane: uniform control flow
  F ep = warp_shfl<ZK_SHFL_UP>(incl, 1), es = warp_shfl<ZK_SHFL_DOWN>(suf, 1);
  if (lane > 0) all_inv = all_inv * ep;
  if (lane < 31) all_inv = all_inv * es;
  return all_inv;
#endif
}
enum { ZK_AFF_KEEP = 0, ZK_AFF_START = 1, ZK_AFF_ADD = 2, ZK_AFF_DBL = 3, ZK_AFF_CANCEL = 4 };
// what the slot's next addition is, and its (never zero) denominator
template <class F> ZK_D uint32_t aff_classify(bool newrun, const Affine<F>& p, const Affine<F>& a, F& d) {
  d = F::one();
  if (newrun) return ZK_AFF_START;
  if (p.is_inf()) return ZK_AFF_KEEP;
  if (a.is_inf()) return ZK_AFF_START;
  F dx = p.x - a.x;
  if (!dx.is_zero()) { d = dx; return ZK_AFF_ADD; }
  if (p.y == a.y) { F y2 = p.y.dbl(); if (!y2.is_zero()) { d = y2; return ZK_AFF_DBL; } }
  return ZK_AFF_CANCEL;
}
template <class F> ZK_D Xyzz<F> aff_to_xyzz(const Affine<F>& a) { return Xyzz<F>::from_affine(a); }

template <class F>
ZK_GLOBAL void k_msm_accumulate_affine(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                       const zk_key_t* __restrict__ skey, const uint32_t* __restrict__ offsets,
                                       const uint32_t* __restrict__ counts, MsmShape s, uint32_t K, uint32_t n_rows,
                                       Affine<F>* __restrict__ acc, F* __restrict__ pre, Xyzz<F>* __restrict__ buckets,
                                       Xyzz<F>* __restrict__ head, Xyzz<F>* __restrict__ tail) {
  const size_t tid = ZK_TID;
  const uint32_t lane = (uint32_t)(tid & 31);
  const uint32_t S = 1u << s.lsS, cpr = s.cap >> s.lsS, cpr32 = cpr >> 5, wpr = cpr32 / K;   // warps per row
  const size_t warp = tid >> 5;
  const bool live = warp < (size_t)n_rows * wpr;
  const uint32_t row = live ? (uint32_t)(warp / wpr) : 0u;
  const uint32_t grp0 = live ? (uint32_t)(warp - (size_t)row * wpr) * K : 0u;
  const uint32_t* off = offsets + (size_t)row * s.nb;
  const uint32_t* cnt = counts + (size_t)row * s.nb;
  const uint32_t total = live ? ZK_LDG(off + s.nb - 1) + ZK_LDG(cnt + s.nb - 1) : 0u;
  const uint32_t gstride = 32u << s.lsS;                           // list entries per group
  // groups of this warp that still have an entry r: a prefix [0, n_r) of its K groups
  auto groups_with = [&](uint32_t r) -> uint32_t {
    if (r >= S || total <= r) return 0u;
    const uint32_t g = (total - r + gstride - 1) / gstride;        // groups of the row with base position + r < total
    return g <= grp0 ? 0u : (g - grp0 < K ? g - grp0 : K);
  };
  const uint32_t* row_list = sorted + (size_t)row * s.cap;
  const zk_key_t* row_keys = skey + (size_t)row * s.cap;
  const size_t slot0 = ((size_t)row * cpr32 + grp0) * 32 + lane;
  F inv = F::one();
  ZK_NOUNROLL for (uint32_t r = 0; r < S; r++) {
    const uint32_t n_r = groups_with(r), n_next = groups_with(r + 1);
    if (n_r == 0) break;
    const bool up = !(r & 1u);
    F prod = F::one();
    // list word of the NEXT slot of the sweep, loaded one iteration early so that its base can be prefetched a full
    // iteration before it is needed (lanes without an entry there prefetch nothing)
    uint32_t e_ahead = 0;
    if (n_r > 1) e_ahead = row_list[(grp0 + (up ? 1u : n_r - 2)) * gstride + (r << 5) + lane];
    ZK_NOUNROLL for (uint32_t j = 0; j < n_r; j++) {
      const uint32_t k = up ? j : n_r - 1 - j;
      const uint32_t at = (grp0 + k) * gstride + (r << 5) + lane;  // chunk-transposed list address of entry r
      const size_t slot = slot0 + (size_t)k * 32;
      const uint32_t chunk = ((grp0 + k) << 5) + lane, pos0 = chunk << s.lsS, pos = pos0 + r;
      const bool active = pos < total;
      const bool next_grp = k < n_next;                            // warp-uniform: this group also takes part in sweep r+1
      const bool next_lane = next_grp && pos + 1 < total;
      // ---- prefetch for the following slot of this sweep
      if (j + 1 < n_r) {
        const size_t slot_n = up ? slot + 32 : slot - 32;
        const uint32_t pos_n = up ? pos + (gstride << 0) : pos - gstride;   // same lane and r, next group: S*32 positions away
        if (pos_n < total) ZK_PREFETCH(bases + (e_ahead & 0x7FFFFFFFu));
        if (j + 2 < n_r) e_ahead = row_list[up ? at + 2 * gstride : at - 2 * gstride];
        if (r) { ZK_PREFETCH(acc + slot_n); ZK_PREFETCH(pre + slot_n); }
      }
      // ---- loads of this slot
      uint32_t key = 0, prevkey = 0, key2 = 0, e = 0, e2 = 0;
      if (active) {
        key = row_keys[at];
        prevkey = r ? row_keys[at - 32] : key;
        e = row_list[at];
        if (next_lane) { key2 = row_keys[at + 32]; e2 = row_list[at + 32]; }
      }
      const bool newrun = r == 0 || key != prevkey;
      const bool next_same = next_lane && key2 == key;             // entry r+1 continues this run: a real addition
      F d = F::one();
      uint32_t mode = ZK_AFF_KEEP;
      Affine<F> p, a;
      if (active) {
        p = bases[e & 0x7FFFFFFFu];
        if (next_same) ZK_PREFETCH(bases + (e2 & 0x7FFFFFFFu));
        if (e >> 31) p.y = p.y.neg();
        if (r) a = acc[slot];
        mode = aff_classify(newrun, p, a, d);
      }
      // ---- finish addition r
      F dinv = inv;
      if (r) {                                                     // sweep 0 only starts sums: nothing to invert
        dinv = F::mul_hot(inv, pre[slot]);
        inv = F::mul_hot(inv, d);
      }
      if (active) {
        const size_t hidx = (size_t)row * cpr + chunk;
        if (newrun && r) {   // the run of `prevkey` ended with the previous entry: whole iff it began inside this chunk
          if (off[prevkey] < pos0) head[hidx] = aff_to_xyzz(a); else buckets[(size_t)row * s.nb + prevkey] = aff_to_xyzz(a);
        }
        if (mode == ZK_AFF_START) a = p;
        else if (mode == ZK_AFF_CANCEL) { a.x = F::zero(); a.y = F::zero(); }
        else if (mode == ZK_AFF_ADD) {
          const F lam = F::mul_hot(p.y - a.y, dinv);
          const F x3 = F::sqr_hot(lam) - a.x - p.x;
          a.y = F::mul_hot(lam, a.x - x3) - a.y;
          a.x = x3;
        } else if (mode == ZK_AFF_DBL) {
          F xx = a.x.sqr();
          const F lam = (xx.dbl() + xx) * dinv;
          const F x3 = lam.sqr() - a.x.dbl();
          a.y = lam * (a.x - x3) - a.y;
          a.x = x3;
        }
        const uint32_t pos1 = pos0 + S < total ? pos0 + S : total;
        if (pos + 1 == pos1) {   // last entry of the chunk: flush the run it belongs to
          const uint32_t st = off[key];
          if (st >= pos0 && st + cnt[key] <= pos1) buckets[(size_t)row * s.nb + key] = aff_to_xyzz(a);
          else if (st <= pos0) head[hidx] = aff_to_xyzz(a);   // the run covering the chunk's first entry
          else tail[hidx] = aff_to_xyzz(a);                    // began inside this chunk, continues in the next
        } else if (mode != ZK_AFF_KEEP) {
          acc[slot] = a;
        }
      }
      // ---- prepare addition r + 1: its denominator joins the prefix products of the next sweep (opposite direction)
      if (next_grp) {
        F d2 = F::one();
        if (next_same) {
          Affine<F> p2 = bases[e2 & 0x7FFFFFFFu];
          if (e2 >> 31) p2.y = p2.y.neg();
          aff_classify(false, p2, a, d2);
        }
        pre[slot] = prod;
        prod = F::mul_hot(prod, d2);
      }
    }
    if (n_next == 0) break;
    inv = coop_inverse(prod, lane);
  }
}

// pass 4b: one thread per (row, bucket): empty buckets become infinity, buckets spread over several chunks are
// summed from the partials those chunks left.
// BOUND = 1: registers capped for one more CTA per SM (G1: 136 -> 128 registers, 4 CTAs; G2: 252 -> 168, 3 CTAs)
// HEAVY buckets (a run crossing more than heavy_span chunks: the bits and small values of a large witness put 10^5 entries -- more
// than 10^4 chunks -- into the bucket of digit 1) are not summed by their one thread, thousands of dependent additions (8.6 ms of a
// 28 ms proof at 2^20 constraints), but cut into SEGMENTS of ZK_HEAVY_SEG chunks and queued: heavy[0] = slots used, then per slot
// four words (bucket id, segment, segments of the bucket, first slot of the bucket).  k_msm_fixup_heavy sums one segment per warp
// (phase 0) and then the segment sums of one bucket per warp (phase 1): ~2 * (8 + 5) dependent additions whatever the bucket size.
// heavy == NULL: no queue.  A bucket that does not fit the queue is summed here after all.
#define ZK_HEAVY_SEG 256u
// fq != NULL (large batches): the cut buckets are only SORTED OUT here -- ids of buckets cut once go to queue 0, of buckets cut more
// often to queue 1 (fq = [n0, n1, -, - | queue 0: q_cap ids | queue 1: q_cap ids]) -- and k_msm_fixup_apply sums them, one thread per
// queued id: every lane of a warp then has an addition to make, and the same number of them.  One thread per bucket summing in
// place ran with 16 of 32 lanes active (ncu, profiles/r02_ncu_full_fixup_reduce_b1024.csv): most buckets are whole or cut once,
// a few twice, and the warp waits for those.
#ifndef ZKFL_EMUL
__device__ __forceinline__ void zk_warp_push(uint32_t* counter, uint32_t* queue, bool pred, uint32_t value) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);       // every lane of the warp arrives here (no early exits before)
  if (!m) return;
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) queue[base + __popc(m & ((1u << lane) - 1u))] = value;
}
#else
static inline void zk_warp_push(uint32_t* counter, uint32_t* queue, bool pred, uint32_t value) {
  if (pred) queue[ZK_ATOMIC_ADD(counter, 1u)] = value;
}
#endif
template <class F, int BOUND>
ZK_GLOBAL ZK_FIX_BOUNDS(F, BOUND) void k_msm_fixup(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts, MsmShape s, uint32_t S,
                           uint32_t chunks_per_row, const Xyzz<F>* __restrict__ head, const Xyzz<F>* __restrict__ tail,
                           Xyzz<F>* __restrict__ buckets, uint32_t heavy_span, uint32_t heavy_cap, uint32_t* __restrict__ heavy,
                           uint32_t* __restrict__ fq, uint32_t q_cap) {
  size_t tid = ZK_TID;
  if (fq) {
    uint32_t cuts = 0;
    if (tid < (size_t)s.B * s.R * s.nb) {
      const uint32_t st = offsets[tid], cnt = counts[tid];
      if (cnt == 0) buckets[tid] = Xyzz<F>::infinity();
      else cuts = (st + cnt - 1) / S - st / S;
    }
    zk_warp_push(fq, fq + 4, cuts == 1, (uint32_t)tid);
    zk_warp_push(fq + 1, fq + 4 + q_cap, cuts > 1, (uint32_t)tid);
    return;
  }
  if (tid >= (size_t)s.B * s.R * s.nb) return;
  size_t row = tid / s.nb;
  uint32_t st = offsets[tid], cnt = counts[tid];
  if (cnt == 0) { buckets[tid] = Xyzz<F>::infinity(); return; }
  uint32_t c0 = st / S, c1 = (st + cnt - 1) / S;
  if (c0 == c1) return;  // written whole by its chunk
  if (heavy && c1 - c0 > heavy_span) {
    const uint32_t nseg = (c1 - c0 + ZK_HEAVY_SEG) / ZK_HEAVY_SEG;          // ceil((c1 - c0 + 1) / SEG)
    const uint32_t slot0 = ZK_ATOMIC_ADD(heavy, nseg);
    if (slot0 + nseg <= heavy_cap) {
      for (uint32_t g = 0; g < nseg; g++) {
        uint32_t* q = heavy + 4 + 4 * (size_t)(slot0 + g);
        q[0] = (uint32_t)tid; q[1] = g; q[2] = nseg; q[3] = slot0;
      }
      return;
    }
    // queue full (the count stays above heavy_cap, which k_msm_fixup_heavy clamps): fall through and sum it here
  }
  const Xyzz<F>* h = head + row * chunks_per_row;
  const Xyzz<F>* t = tail + row * chunks_per_row;
  Xyzz<F> acc = (st > c0 * S) ? t[c0] : h[c0];
  ZK_NOUNROLL for (uint32_t ch = c0 + 1; ch <= c1; ch++) {
    if (BOUND == 2) xyzz_add_hot(acc, h[ch]);      // few rows: a chain of dependent additions, products overlapped
    else xyzz_add(acc, h[ch]);
  }
  buckets[tid] = acc;
}
// thread i of queue `which`: the bucket's partial sums, tail (or head) of its first chunk plus the heads of the following ones
template <class F, int BOUND>
ZK_GLOBAL ZK_FIX_BOUNDS(F, BOUND) void k_msm_fixup_apply(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts, MsmShape s,
                           uint32_t S, uint32_t chunks_per_row, const Xyzz<F>* __restrict__ head, const Xyzz<F>* __restrict__ tail,
                           Xyzz<F>* __restrict__ buckets, const uint32_t* __restrict__ fq, uint32_t q_cap, uint32_t which) {
  const size_t i = ZK_TID;
  if (i >= fq[which]) return;
  const uint32_t id = fq[4 + (size_t)which * q_cap + i];
  const size_t row = id / s.nb;
  const uint32_t st = offsets[id], cnt = counts[id], c0 = st / S, c1 = (st + cnt - 1) / S;
  const Xyzz<F>* h = head + row * chunks_per_row;
  const Xyzz<F>* t = tail + row * chunks_per_row;
  Xyzz<F> acc = (st > c0 * S) ? t[c0] : h[c0];
  for (uint32_t ch = c0 + 1; ch <= c1; ch++) xyzz_add(acc, h[ch]);
  buckets[id] = acc;
}
#ifndef ZKFL_EMUL
// one WARP per queue slot.  phase 0: lane l sums the partials of the slot's segment, chunks c0 + seg * SEG + l, + 32, ...; the 32 lane
// sums meet in a shuffle tree (all lanes run the same additions, lane 0 holds the result) -> hsum[slot].  phase 1 (slots with
// segment 0 only): the same over the bucket's segment sums -> buckets[id].
template <class F>
static __global__ void k_msm_fixup_heavy(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts, MsmShape s, uint32_t S,
                                         uint32_t chunks_per_row, const Xyzz<F>* __restrict__ head, const Xyzz<F>* __restrict__ tail,
                                         Xyzz<F>* __restrict__ buckets, uint32_t heavy_cap, const uint32_t* __restrict__ heavy,
                                         Xyzz<F>* __restrict__ hsum, int phase) {
  const uint32_t warp = (uint32_t)(ZK_TID >> 5), lane = threadIdx.x & 31u;
  // slots are handed out in whole buckets: the last bucket that asked may not have fitted; every slot below its slot0 is valid.
  // A slot is valid iff its record says so: records are written only by buckets that fitted, stale ones are cleared by the host memset.
  if (warp >= heavy_cap) return;
  const uint32_t* q = heavy + 4 + 4 * (size_t)warp;
  const uint32_t id = q[0], seg = q[1], nseg = q[2], slot0 = q[3];
  if (nseg == 0) return;                                    // unused slot
  if (phase == 1 && seg != 0) return;
  // ONE addition site (inlined products, see xyzz_add_hot): first the lane's strided partials -- every lane makes the same number of
  // trips, the ones past the end add infinity -- then the five rounds of the shuffle tree
  const Xyzz<F>* src = hsum + slot0;       // phase 1: the bucket's segment sums
  const Xyzz<F>* first_src = nullptr;      // phase 0, segment 0 of a run that starts inside its first chunk: that chunk's TAIL partial
  uint32_t n_items = nseg;
  if (phase == 0) {
    const size_t row = id / s.nb;
    const uint32_t st = offsets[id], cnt = counts[id], c0 = st / S, c1 = (st + cnt - 1) / S;
    const uint32_t lo = c0 + seg * ZK_HEAVY_SEG, hi = lo + ZK_HEAVY_SEG - 1 < c1 ? lo + ZK_HEAVY_SEG - 1 : c1;
    src = head + row * chunks_per_row + lo;
    n_items = hi - lo + 1;
    if (seg == 0 && st > c0 * S) first_src = tail + row * chunks_per_row + c0;
  }
  Xyzz<F> acc = Xyzz<F>::infinity();
  const uint32_t trips = (n_items + 31) / 32;                 // the same for every lane of the warp
  ZK_NOUNROLL for (uint32_t i = 0; i < trips + 5; i++) {
    Xyzz<F> o;
    if (i < trips) {
      const uint32_t g = i * 32 + lane;
      if (g < n_items) o = (g == 0 && first_src) ? *first_src : src[g];
      else o = Xyzz<F>::infinity();
    } else {
      const uint32_t off = 16u >> (i - trips);
      o.X = warp_shfl<ZK_SHFL_DOWN>(acc.X, off); o.Y = warp_shfl<ZK_SHFL_DOWN>(acc.Y, off);
      o.ZZ = warp_shfl<ZK_SHFL_DOWN>(acc.ZZ, off); o.ZZZ = warp_shfl<ZK_SHFL_DOWN>(acc.ZZZ, off);
    }
    xyzz_add_hot(acc, o);
  }
  if (lane == 0) { if (phase == 0) hsum[warp] = acc; else buckets[id] = acc; }
}
#endif
// pass 5: bucket reduction S = sum_k (k+1) * X[k] over the nb buckets of a row, as a three-level tree so that the
// serial depth is ~ 2*L1 + 2*L2 + 5*N2 additions instead of 2*sqrt(nb) + 3*sqrt(nb).
// Level kernel: chunk t of L consecutive elements -> R_t = sum X, T_t = sum j * X[t*L + j] (zero-based local weights).
// With Z(X) = sum_k k * X[k]:  Z(X) = sum_t T_t + L * Z(R)  and  S = Z(X) + sum(X) = Z(X) + sum(R).
// MEASURED AND REJECTED in round 2 (B200, 1024 sgd_verified proofs per step, step 288 ms with these kernels): (a) the fix-up fused
// into this level (partials added straight into the running sum) with the products inlined at four call sites: 336 ms (the kernel
// falls out of the instruction cache); (b) the same around ONE inlined addition fed by selects: 295 ms; (c) plus the second
// accumulator parked in shared memory to reach 3 / 4 CTAs per SM: 370 / 326 ms.  Serialised (ncu) the fused level ran at ~0.72 ns per
// addition against 0.75 ns for this call-based form: all variants sit at ~30 % of the multiplier rate because a thread is one long
// chain of dependent additions at 2 warps per scheduler; concentrating the fix-up work in this low-occupancy kernel only made it
// worse than leaving it in the massively parallel k_msm_fixup.  What would pay is fewer additions, not a different packaging.
template <class F, int BOUND>
ZK_GLOBAL ZK_LVL_BOUNDS(F, BOUND) void k_reduce_level(const Xyzz<F>* __restrict__ in, size_t rows, uint32_t N, uint32_t L, Xyzz<F>* __restrict__ R,
                                   Xyzz<F>* __restrict__ T) {
  size_t tid = ZK_TID;
  uint32_t nchunk = N / L;
  if (tid >= rows * nchunk) return;
  size_t row = tid / nchunk;
  uint32_t ch = (uint32_t)(tid % nchunk);
  const Xyzz<F>* x = in + row * N + (size_t)ch * L;
  Xyzz<F> run = Xyzz<F>::infinity(), acc = Xyzz<F>::infinity();
  for (int j = (int)L - 1; j >= 1; j--) {
    xyzz_add(run, x[j]);
    xyzz_add(acc, run);
  }
  xyzz_add(run, x[0]);
  R[tid] = run;
  if (T) T[tid] = acc;
}
// final: per row, from the level-2 outputs (N2 entries each): R2/T2 = level 2 of R1, RT = chunk sums of T1.
//   Z(R1) = sum(T2) + L2 * Z(R2);  Z(X) = sum(T1) + L1 * Z(R1) = sum(RT) + L1 * Z(R1);  S = Z(X) + sum(R2)
template <class F>
ZK_GLOBAL void k_reduce_final(const Xyzz<F>* __restrict__ R2, const Xyzz<F>* __restrict__ T2, const Xyzz<F>* __restrict__ RT,
                              size_t rows, uint32_t N2, uint32_t L1, uint32_t L2, Xyzz<F>* __restrict__ out) {
  size_t row = ZK_TID;
  if (row >= rows) return;
  const Xyzz<F>* r2 = R2 + row * N2;
  const Xyzz<F>* t2 = T2 + row * N2;
  const Xyzz<F>* rt = RT + row * N2;
  Xyzz<F> run = Xyzz<F>::infinity(), z = Xyzz<F>::infinity(), sum_r = Xyzz<F>::infinity(), sum_t2 = Xyzz<F>::infinity(),
          sum_t1 = Xyzz<F>::infinity();
  for (int t = (int)N2 - 1; t >= 1; t--) { xyzz_add(run, r2[t]); xyzz_add(z, run); }   // Z(R2)
  for (uint32_t t = 0; t < N2; t++) { xyzz_add(sum_r, r2[t]); xyzz_add(sum_t2, t2[t]); xyzz_add(sum_t1, rt[t]); }
  for (uint32_t l = L2; l > 1; l >>= 1) z = xyzz_dbl(z);
  xyzz_add(z, sum_t2);                                                                    // Z(R1)
  for (uint32_t l = L1; l > 1; l >>= 1) z = xyzz_dbl(z);
  xyzz_add(z, sum_t1);                                                                    // Z(X)
  xyzz_add(z, sum_r);                                                                     // + sum(X)
  out[row] = z;
}
// pass 5, LATENCY variant (few rows: single proofs, the per-rank share of a split proof).  With one row the tree above is a
// serial chain of ~2*L1 + 2*L2 + 5*N2 additions (288 for 2^15 buckets: 9-12 ms).  Here the weights are taken bit by bit:
//   S = sum_k (k+1) X[k] = Y_all + sum_b 2^b Y_b,   Y_b = sum of the X[k] whose index has bit b set,
// so everything is PLAIN sums, done as a fan-in-L tree (L = 8: three index bits per level).  One launch per level; thread =
// (part, row, output element): part 0 sums a chunk of the main array (-> next main array), parts 1..n_pool sum a chunk of a
// pending bit array, the last lgL parts sum the chunk elements whose local index has bit b set (-> new pending arrays).
// Every thread adds at most L points; depth = L per level + 2 per bit in the final Horner (~70 additions for 2^15 buckets),
// about 3*nb additions of work per row instead of 2*nb -- irrelevant at this size, the GPU is otherwise idle.
// pool layout: [array][row][element]; arrays are in bit order (level 0 creates bits 0..lgL-1, and so on).
template <class F>
ZK_GLOBAL void k_reduce_bits_level(const Xyzz<F>* __restrict__ main_in, const Xyzz<F>* __restrict__ pool_in, uint32_t n_pool_in,
                                   size_t rows, uint32_t N_in, uint32_t lgL, Xyzz<F>* __restrict__ main_out,
                                   Xyzz<F>* __restrict__ pool_out) {
  const uint32_t N_out = N_in >> lgL, L = 1u << lgL;
  const size_t per_part = rows * N_out, tid = ZK_TID;
  if (tid >= per_part * (1 + n_pool_in + lgL)) return;
  const uint32_t part = (uint32_t)(tid / per_part);
  const size_t rem = tid % per_part, row = rem / N_out;
  const uint32_t t = (uint32_t)(rem % N_out);
  const Xyzz<F>* src;
  Xyzz<F>* dst;
  uint32_t bit = 0xFFFFFFFFu;                       // no filter: every element of the chunk
  if (part == 0) {
    src = main_in + row * N_in + (size_t)t * L;
    dst = main_out + row * N_out + t;
  } else if (part <= n_pool_in) {
    const size_t a = part - 1;
    src = pool_in + (a * rows + row) * N_in + (size_t)t * L;
    dst = pool_out + (a * rows + row) * N_out + t;
  } else {
    bit = part - 1 - n_pool_in;
    src = main_in + row * N_in + (size_t)t * L;
    dst = pool_out + ((size_t)(n_pool_in + bit) * rows + row) * N_out + t;
  }
  Xyzz<F> acc = Xyzz<F>::infinity();
  for (uint32_t j = 0; j < L; j++)
    if (bit == 0xFFFFFFFFu || ((j >> bit) & 1u)) xyzz_add_hot(acc, src[j]);   // latency-bound: overlapping products
  *dst = acc;
}
// main: [rows] (Y_all), pool: [n_bits][rows] (Y_b): out[row] = Y_all + sum_b 2^b Y_b by Horner from the top bit
template <class F>
ZK_GLOBAL void k_reduce_bits_final(const Xyzz<F>* __restrict__ main_in, const Xyzz<F>* __restrict__ pool, uint32_t n_bits, size_t rows,
                                   Xyzz<F>* __restrict__ out) {
  size_t row = ZK_TID;
  if (row >= rows) return;
  Xyzz<F> z = Xyzz<F>::infinity();
  ZK_NOUNROLL for (int b = (int)n_bits - 1; b >= -1; b--) {      // b = -1: the unweighted term (one addition site for all of them)
    if (b >= 0) z = xyzz_dbl_hot(z);
    xyzz_add_hot(z, b >= 0 ? pool[(size_t)b * rows + row] : main_in[row]);
  }
  out[row] = z;
}
// pass 6: Horner over the windows, one thread per proof: out[b] = sum_j 2^(c*j) * win[b][j]
template <class F>
ZK_GLOBAL void k_msm_combine(const Xyzz<F>* __restrict__ win, MsmShape s, Xyzz<F>* __restrict__ out) {
  size_t b = ZK_TID;
  if (b >= s.B) return;
  Xyzz<F> acc = win[b * s.R + (s.R - 1)];
  for (int j = (int)s.R - 2; j >= 0; j--) {
    for (uint32_t q = 0; q < s.c; q++) acc = xyzz_dbl(acc);
    xyzz_add(acc, win[b * s.R + j]);
  }
  out[b] = acc;
}

// ================================================================================ per-zkey precomputation
// table[j*m + i] = 2^(c*j) * P_i for j < W (affine Montgomery): the bases are per-circuit constants shared by every
// proof, so the window shifts are paid once at zkey load instead of c doublings per window per proof.
template <class F>
ZK_GLOBAL void k_precompute_windows(const Affine<F>* __restrict__ bases, uint32_t m, uint32_t c, uint32_t W,
                                    Affine<F>* __restrict__ table) {
  size_t i = ZK_TID;
  if (i >= m) return;
  Affine<F> p = bases[i];
  table[i] = p;
  Xyzz<F> q = Xyzz<F>::from_affine(p);
  for (uint32_t j = 1; j < W; j++) {
    for (uint32_t k = 0; k < c; k++) q = xyzz_dbl(q);
    Affine<F> a = xyzz_to_affine(q);
    table[(size_t)j * m + i] = a;
    q = Xyzz<F>::from_affine(a);
  }
}
// fixed-base byte-window table: tab[j*256 + d] = (d << 8j) * base, j < 32 (entry d = 0 is infinity)
template <class F>
ZK_GLOBAL void k_fixed_base_table(Affine<F> base, Affine<F>* __restrict__ tab) {
  size_t tid = ZK_TID;
  if (tid >= 32 * 256) return;
  uint32_t j = (uint32_t)(tid >> 8), d = (uint32_t)(tid & 255);
  uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  k[j >> 2] = d << (8 * (j & 3));
  if (j == 31 && d >= 64) { Affine<F> z; z.x = F::zero(); z.y = F::zero(); tab[tid] = z; return; }  // beyond 254 bits
  tab[tid] = xyzz_to_affine(xyzz_scalar_mul(Xyzz<F>::from_affine(base), k));
}
// k * base from the byte-window table: 32 mixed adds
template <class F>
ZK_D Xyzz<F> fixed_base_mul(const Affine<F>* __restrict__ tab, const uint32_t* k) {
  Xyzz<F> acc = Xyzz<F>::infinity();
  for (uint32_t j = 0; j < 32; j++) {
    uint32_t d = (k[j >> 2] >> (8 * (j & 3))) & 255u;
    if (d) xyzz_madd(acc, tab[j * 256 + d], false);
  }
  return acc;
}

// ================================================================================ misc
// out[i] = k_i * G as Montgomery affine (zkey point layout): `groth16 setup`'s scalar multiplications
template <class F>
ZK_GLOBAL void k_gen_mul(Affine<F> gen, const Fr* __restrict__ scalars, size_t n, Affine<F>* __restrict__ out) {
  size_t i = ZK_TID;
  if (i >= n) return;
  Fr k = scalars[i];
  out[i] = xyzz_to_affine(xyzz_scalar_mul(Xyzz<F>::from_affine(gen), k.v));
}
// out[i] = k * P_i for ONE scalar k (Montgomery affine in and out): `snarkjs zkey contribute` rescales the C and H sections by
// 1/d and delta by d (tests/full_system_simulation.mjs:726-731)
template <class F>
ZK_GLOBAL void k_point_scale(const Affine<F>* __restrict__ pts, Fr k, size_t n, Affine<F>* __restrict__ out) {
  size_t i = ZK_TID;
  if (i >= n) return;
  out[i] = xyzz_to_affine(xyzz_scalar_mul(Xyzz<F>::from_affine(pts[i]), k.v));
}
// a single XYZZ result -> affine canonical bytes (standalone MSM API)
template <class F>
ZK_GLOBAL void k_to_affine_canonical(const Xyzz<F>* __restrict__ in, size_t n, Affine<F>* __restrict__ out) {
  size_t i = ZK_TID;
  if (i >= n) return;
  Affine<F> a = xyzz_to_affine(in[i]);
  a.x = a.x.from_mont();
  a.y = a.y.from_mont();
  out[i] = a;
}
// affine canonical bytes -> XYZZ Montgomery, summing `nparts` partial results per output (multi-GPU split MSM:
// every rank contributes one partial per MSM; the group law is not an NCCL reduction op, so "reduce" = gather + add)
template <class F>
ZK_GLOBAL void k_sum_partials(const Affine<F>* __restrict__ parts, uint32_t nparts, size_t part_stride, size_t elem_stride,
                              size_t n, Xyzz<F>* __restrict__ out) {
  size_t i = ZK_TID;
  if (i >= n) return;
  Xyzz<F> acc = Xyzz<F>::infinity();
  for (uint32_t p = 0; p < nparts; p++) {
    Affine<F> a = parts[p * part_stride + i * elem_stride];
    if (!a.is_inf()) { a.x = to_mont_any(a.x); a.y = to_mont_any(a.y); }
    xyzz_madd(acc, a, false);
  }
  out[i] = acc;
}
#ifdef ZK_K_MSM_SORT
// marks the points outside [lo, hi) (and those already skipped) so a rank only sorts its own range
ZK_GLOBAL void k_range_mask(const uint8_t* __restrict__ base_skip, uint32_t m, uint32_t lo, uint32_t hi, uint8_t* __restrict__ out) {
  size_t i = ZK_TID;
  if (i >= m) return;
  out[i] = (i < lo || i >= hi || (base_skip && base_skip[i])) ? 1 : 0;
}

#endif  // ZK_K_MSM_SORT
}  // namespace zk
