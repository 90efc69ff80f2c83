// sm_100a kernels of the Groth16 proving path.
//
// Data layout in HBM: everything that has one value per client proof is stored batch-minor,
// `x[element][b]` with b < B the proof index ("SoA across proofs").  A warp therefore touches
// 32 consecutive 32-byte field elements (1 KiB, fully coalesced) for ANY element stride, which
// makes every NTT stage, the sparse A.w/B.w products and the witness program coalesced without
// shared-memory staging; per-element constants (twiddles, matrix coefficients, round constants)
// are warp-uniform broadcast loads.  Bases (zkey points) are shared by all proofs of a batch and
// stay L2-resident (a few MB per circuit).
//
// All kernels are sync-free one-thread-per-item kernels (which is also what lets the CPU test-suite run the same
// sources as a host emulation): the bound is the integer pipe (fused IMAD.WIDE multiply-accumulates) for the group
// law, the Montgomery products and -- at the batch sizes used -- even the NTT passes; tensor cores do not apply
// (no dense contraction anywhere on the path). Balance comes from the work decomposition (fixed-size chunks of the
// bucket-sorted lists, dependency levels of the witness program), not from intra-CTA cooperation.
// Non-template kernels are grouped in sections (ZK_K_WITNESS, ZK_K_MSM_SORT, ZK_K_FIN, ZK_K_VERIFY, ZK_K_BENCH): every .cu file of
// the library defines the sections it launches before including this header, so each kernel is compiled exactly once.
#pragma once
#include "bn254.cuh"
#include "pairing.cuh"

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
#define ZK_RED_BOUNDS __launch_bounds__(128)     // bucket-reduction kernels: 128-thread CTAs, registers as needed (no spills)
#else
#define ZK_ACC_BOUNDS(F)
#define ZK_RED_BOUNDS
#endif

// ================================================================================ layout helpers
#ifdef ZK_K_WITNESS
// host layout [b][e] (what .wtns / the C ABI use)  ->  device layout [e][b]
ZK_GLOBAL void k_aos_to_soa(const Fr* __restrict__ src, Fr* __restrict__ dst, uint32_t n_elem, uint32_t B,
                            uint32_t dst_elem_off) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)n_elem * B) return;
  uint32_t e = (uint32_t)(tid / B), b = (uint32_t)(tid % B);
  dst[(size_t)(e + dst_elem_off) * B + b] = src[(size_t)b * n_elem + e];
}
ZK_GLOBAL void k_soa_to_aos(const Fr* __restrict__ src, Fr* __restrict__ dst, uint32_t n_elem, uint32_t B) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)n_elem * B) return;
  uint32_t e = (uint32_t)(tid / B), b = (uint32_t)(tid % B);
  dst[(size_t)b * n_elem + e] = src[(size_t)e * B + b];
}

// selected wires of the device witness [n_wires][B] -> host layout [b][k] (used to read commitments out of a program run)
ZK_GLOBAL void k_gather_wires(const Fr* __restrict__ w, const uint32_t* __restrict__ wires, uint32_t n_sel, uint32_t B,
                              Fr* __restrict__ out) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)n_sel * B) return;
  uint32_t k = (uint32_t)(tid / B), b = (uint32_t)(tid % B);
  out[(size_t)b * n_sel + k] = w[(size_t)ZK_LDG(wires + k) * B + b];
}

#endif  // ZK_K_WITNESS
// ================================================================================ W1: batched witness evaluator
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

#ifdef ZK_K_WITNESS
ZK_D Fr lc_eval(const ProgramDev& p, uint32_t k, const Fr* __restrict__ w, uint32_t B, uint32_t b) {
  Fr acc = Fr::zero();
  uint32_t e = ZK_LDG(p.lc_off + k + 1);
  for (uint32_t i = ZK_LDG(p.lc_off + k); i < e; i++)
    acc = acc + p.lc_coef[i] * w[(size_t)ZK_LDG(p.lc_wire + i) * B + b];
  return acc;
}

// w[0][b] = 1 for every client instance
ZK_GLOBAL void k_witness_init(Fr* __restrict__ w, uint32_t B) {
  uint32_t b = (uint32_t)ZK_TID;
  if (b >= B) return;
  Fr one_c = Fr::zero(); one_c.v[0] = 1;
  w[b] = one_c;
}
// One dependency level of the witness program: thread = (op of the level, client instance), clients minor, so a warp
// runs ONE op for 32 clients in lock-step (no divergence) and all its loads/stores are coalesced. The compiler
// (circuits/builder.py) sorts ops by level; ops inside a level are independent of each other.
// w: canonical witness [n_wires][B]; rows 1..n_inputs hold the inputs.
ZK_GLOBAL void k_witness_level(ProgramDev p, Fr* __restrict__ w, uint32_t B, uint32_t op_lo, uint32_t op_hi) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)(op_hi - op_lo) * B) return;
  const uint32_t o = op_lo + (uint32_t)(tid / B), b = (uint32_t)(tid % B);
  const uint32_t* op = p.ops + 5 * (size_t)o;
  const uint32_t code = ZK_LDG(op), dst = ZK_LDG(op + 1), a = ZK_LDG(op + 2), bb = ZK_LDG(op + 3), c = ZK_LDG(op + 4);
  if (code == 1) {
    w[(size_t)dst * B + b] = lc_eval(p, a, w, B, b).from_mont();
  } else if (code == 2) {
    Fr v = lc_eval(p, a, w, B, b) * lc_eval(p, bb, w, B, b);
    if (c != 0xFFFFFFFFu) v = v + lc_eval(p, c, w, B, b);
    w[(size_t)dst * B + b] = v.from_mont();
  } else if (code == 3) {
    Fr v = lc_eval(p, a, w, B, b).from_mont();
    for (uint32_t i = 0; i < bb; i++) {
      Fr bit = Fr::zero();
      bit.v[0] = (v.v[i >> 5] >> (i & 31)) & 1u;
      w[(size_t)(dst + i) * B + b] = bit;
    }
  } else if (code == 4) {
    const uint32_t t = a;
    const PoseidonDev K = p.pk[t];
    Fr st[17], nx[17];
    st[0] = Fr::zero();
    for (uint32_t i = 1; i < t; i++) st[i] = w[(size_t)ZK_LDG(p.pos_in + bb + i - 1) * B + b].to_mont();
    size_t k = dst;
    for (uint32_t r = 0; r < K.rounds; r++) {
      for (uint32_t i = 0; i < t; i++) st[i] = st[i] + K.C[r * t + i];
      const uint32_t lanes = (r < 4 || r >= 4 + K.rp) ? t : 1;
      for (uint32_t i = 0; i < lanes; i++) {
        Fr x2 = st[i].sqr(), x4 = x2.sqr(), x5 = x4 * st[i];
        w[k * B + b] = x2.from_mont();
        w[(k + 1) * B + b] = x4.from_mont();
        w[(k + 2) * B + b] = x5.from_mont();
        k += 3;
        st[i] = x5;
      }
      for (uint32_t i = 0; i < t; i++) {
        Fr acc = Fr::zero();
        for (uint32_t j = 0; j < t; j++) acc = acc + K.M[i * t + j] * st[j];
        nx[i] = acc;
      }
      for (uint32_t i = 0; i < t; i++) st[i] = nx[i];
    }
    w[k * B + b] = st[0].from_mont();
  }
}

#ifndef ZKFL_EMUL
// LATENCY variant of a witness level (few client instances): one WARP per (op, client).  A Poseidon permutation is a serial
// chain of rounds; with one thread per instance a round costs t^2 + 6 products (MDS row by row, S-box, canonical copies of the
// three S-box wires), ~1 ms per t = 6 permutation and 5 ms for the 7 dependency levels of sgd_verified.  Here lane i owns state
// element i: a round is t + 6 products per lane (its MDS row over shuffled state, its own S-box).  Every lane runs the same
// instruction stream (lanes >= t mirror element t-1, S-box results are selected, stores are predicated): no divergent calls.
// Other ops are evaluated redundantly by all lanes (same value, same address).  Wire numbering is identical to k_witness_level.
ZK_GLOBAL void k_witness_level_coop(ProgramDev p, Fr* __restrict__ w, uint32_t B, uint32_t op_lo, uint32_t op_hi) {
  const size_t gw = ZK_TID >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (gw >= (size_t)(op_hi - op_lo) * B) return;      // a whole warp at a time
  const uint32_t o = op_lo + (uint32_t)(gw / B), b = (uint32_t)(gw % B);
  const uint32_t* op = p.ops + 5 * (size_t)o;
  const uint32_t code = ZK_LDG(op), dst = ZK_LDG(op + 1), a = ZK_LDG(op + 2), bb = ZK_LDG(op + 3), c = ZK_LDG(op + 4);
  if (code == 1) {
    w[(size_t)dst * B + b] = lc_eval(p, a, w, B, b).from_mont();
  } else if (code == 2) {
    Fr v = lc_eval(p, a, w, B, b) * lc_eval(p, bb, w, B, b);
    if (c != 0xFFFFFFFFu) v = v + lc_eval(p, c, w, B, b);
    w[(size_t)dst * B + b] = v.from_mont();
  } else if (code == 3) {
    Fr v = lc_eval(p, a, w, B, b).from_mont();
    for (uint32_t i = lane; i < bb; i += 32) {      // the bits of the value: one per lane
      Fr bit = Fr::zero();
      bit.v[0] = (v.v[i >> 5] >> (i & 31)) & 1u;
      w[(size_t)(dst + i) * B + b] = bit;
    }
  } else if (code == 4) {
    const uint32_t t = a;
    const PoseidonDev K = p.pk[t];
    const uint32_t li = lane < t ? lane : t - 1;
    Fr st = Fr::zero();
    if (li) st = w[(size_t)ZK_LDG(p.pos_in + bb + li - 1) * B + b];
    st = st.to_mont();                                // (0 stays 0)
    size_t k = dst;
    ZK_NOUNROLL for (uint32_t r = 0; r < K.rounds; r++) {
      st = st + K.C[r * t + li];
      const bool full = r < 4 || r >= 4 + K.rp;
      const Fr x2 = st.sqr(), x4 = x2.sqr(), x5 = x4 * st;
      const Fr c2 = x2.from_mont(), c4 = x4.from_mont(), c5 = x5.from_mont();
      if (full ? lane < t : lane == 0) {
        const size_t kk = k + (full ? 3 * (size_t)li : 0);
        w[kk * B + b] = c2;
        w[(kk + 1) * B + b] = c4;
        w[(kk + 2) * B + b] = c5;
      }
      if (full || li == 0) st = x5;
      k += full ? 3 * (size_t)t : 3;
      Fr acc = Fr::zero();
      ZK_NOUNROLL for (uint32_t j = 0; j < t; j++) {
        const Fr sj = warp_shfl<ZK_SHFL_IDX>(st, j);
        acc = acc + K.M[li * t + j] * sj;
      }
      st = acc;
    }
    const Fr out = st.from_mont();
    if (lane == 0) w[k * B + b] = out;
  }
}
#endif

#endif  // ZK_K_WITNESS
// ================================================================================ K1: sparse A.w, B.w, C = A o B
struct CsrDev {
  const uint32_t* row_off;  // n_rows + 1
  const uint32_t* wire;
  const Fr* coef;           // coef * R^2 (the bytes zkey section 4 stores)
};
#ifdef ZK_K_WITNESS
ZK_D Fr csr_row(const CsrDev& m, uint32_t row, const Fr* __restrict__ w, uint32_t B, uint32_t b) {
  Fr acc = Fr::zero();
  uint32_t e = ZK_LDG(m.row_off + row + 1);
  for (uint32_t i = ZK_LDG(m.row_off + row); i < e; i++)
    acc = acc + m.coef[i] * w[(size_t)ZK_LDG(m.wire + i) * B + b];
  return acc;
}
// abc: [3][n][B] Montgomery (A evals, B evals, C = A*B) over the constraint domain
ZK_GLOBAL void k_build_abc(CsrDev A, CsrDev Bm, const Fr* __restrict__ w, Fr* __restrict__ abc, uint32_t n, uint32_t B) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)n * B) return;
  uint32_t row = (uint32_t)(tid / B), b = (uint32_t)(tid % B);
  Fr a = csr_row(A, row, w, B, b), bv = csr_row(Bm, row, w, B, b);
  abc[tid] = a;
  abc[(size_t)n * B + tid] = bv;
  abc[2 * (size_t)n * B + tid] = a * bv;
}
// constraint check (what a failing `===` is for circom): first violated row per client, or 0xFFFFFFFF
ZK_GLOBAL void k_r1cs_check(CsrDev A, CsrDev Bm, CsrDev C, const Fr* __restrict__ w, uint32_t n_rows, uint32_t B,
                            uint32_t* __restrict__ first_bad) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)n_rows * B) return;
  uint32_t row = (uint32_t)(tid / B), b = (uint32_t)(tid % B);
  Fr a = csr_row(A, row, w, B, b), bv = csr_row(Bm, row, w, B, b), c = csr_row(C, row, w, B, b);
  // a, bv, c are Montgomery(value): a*bv = Mont(product), compare in Montgomery form
  if (!((a * bv) == c)) ZK_ATOMIC_MIN(first_bad + b, row);
}

// well-formedness of a witness handed in from outside (`groth16 prove <zkey> <wtns>`): every element reduced mod r, wire 0 == 1.
// flags: bit 0 = some element >= r, bit 1 = some instance has w[0] != 1
ZK_GLOBAL void k_wtns_validate(const Fr* __restrict__ w, uint32_t n_wires, uint32_t B, uint32_t* __restrict__ flags) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)n_wires * B) return;
  const Fr v = w[tid];
  bool lt = false, decided = false;
  ZK_UNROLL for (int i = 7; i >= 0; i--) {
    if (!decided && v.v[i] != FrP::mod(i)) { lt = v.v[i] < FrP::mod(i); decided = true; }
  }
  uint32_t f = lt ? 0u : 1u;
  if (tid < B) { uint32_t o = v.v[0] ^ 1u; ZK_UNROLL for (int i = 1; i < 8; i++) o |= v.v[i]; if (o) f |= 2u; }
  if (f) ZK_ATOMIC_OR(flags, f);
}

// ================================================================================ K2-K5: H polynomial
// data: [n_poly][n][B]; one radix-2 stage. dif=1: Gentleman-Sande (natural in -> bit-reversed out after
// all stages, used for the inverse transform), dif=0: Cooley-Tukey (bit-reversed in -> natural out).
// tw: n/2 powers of the (inverse) root, Montgomery.
ZK_GLOBAL void k_ntt_stage(Fr* __restrict__ data, const Fr* __restrict__ tw, uint32_t n, uint32_t B, uint32_t n_poly,
                           uint32_t half, int dif) {
  size_t tid = ZK_TID;
  size_t per_poly = (size_t)(n / 2) * B;
  if (tid >= per_poly * n_poly) return;
  uint32_t poly = (uint32_t)(tid / per_poly);
  size_t rem = tid % per_poly;
  uint32_t pr = (uint32_t)(rem / B), b = (uint32_t)(rem % B);
  uint32_t blk = pr / half, k = pr % half;
  size_t i0 = (size_t)blk * 2 * half + k, i1 = i0 + half;
  Fr* base = data + (size_t)poly * n * B;
  Fr u = base[i0 * B + b], v = base[i1 * B + b];
  Fr wv = tw[(size_t)k * (n / 2 / half)];
  if (dif) {
    base[i0 * B + b] = u + v;
    base[i1 * B + b] = (u - v) * wv;
  } else {
    v = v * wv;
    base[i0 * B + b] = u + v;
    base[i1 * B + b] = u - v;
  }
}
// K fused radix-2 stages in registers (K = 1, 2, 3): a thread owns the 2^K elements that form a closed butterfly
// network for those stages, so a pass over HBM does K stages instead of one (14 stages = 5 passes at n = 2^14).
// dif = 1: stages with half = h, h/2, ... (h = `half`, the first stage's half); elements k + r*(h >> (K-1)).
// dif = 0: stages with half = h, 2h, ...; elements k + r*h inside a block of size h << K.
// scale (may be null, dif only): multiply element at position p by scale[p] on the way out (fuses k_scale_rows into
// the last inverse pass).
template <int K>
ZK_GLOBAL void k_ntt_radix(Fr* __restrict__ data, const Fr* __restrict__ tw, const Fr* __restrict__ scale, uint32_t n, uint32_t B,
                           uint32_t n_poly, uint32_t half, int dif) {
  constexpr uint32_t R = 1u << K;
  size_t tid = ZK_TID;
  size_t per_poly = (size_t)(n >> K) * B;
  if (tid >= per_poly * n_poly) return;
  uint32_t poly = (uint32_t)(tid / per_poly);
  size_t rem = tid % per_poly;
  uint32_t g = (uint32_t)(rem / B), b = (uint32_t)(rem % B);
  Fr* base = data + (size_t)poly * n * B;
  Fr x[R];
  if (dif) {
    const uint32_t step = half >> (K - 1);          // distance between the thread's elements
    const uint32_t blk = g / step, k = g % step;     // block of size 2*half
    const size_t i0 = (size_t)blk * 2 * half + k;
    ZK_UNROLL for (uint32_t r = 0; r < R; r++) x[r] = base[(i0 + (size_t)r * step) * B + b];
    ZK_UNROLL for (int s = 0; s < K; s++) {
      const uint32_t h = half >> s;                   // this stage's half, in elements
      const uint32_t hr = R >> (s + 1);               // ... in units of `step`
      ZK_UNROLL for (uint32_t r = 0; r < R; r++) {
        if ((r / hr) & 1) continue;                   // r is the upper element of its pair
        const uint32_t e = k + (r % hr) * step;       // exponent inside the stage's block of size 2h
        Fr u = x[r], v = x[r + hr];
        x[r] = u + v;
        x[r + hr] = (u - v) * tw[(size_t)e * (n / 2 / h)];
      }
    }
    ZK_UNROLL for (uint32_t r = 0; r < R; r++) {
      size_t p = i0 + (size_t)r * step;
      base[p * B + b] = scale ? x[r] * scale[p] : x[r];
    }
  } else {
    const uint32_t blk = g / half, k = g % half;     // block of size half << K
    const size_t i0 = (size_t)blk * ((size_t)half << K) + k;
    ZK_UNROLL for (uint32_t r = 0; r < R; r++) x[r] = base[(i0 + (size_t)r * half) * B + b];
    ZK_UNROLL for (int s = 0; s < K; s++) {
      const uint32_t h = half << s;
      const uint32_t hr = 1u << s;                    // pair distance in units of `half`
      ZK_UNROLL for (uint32_t r = 0; r < R; r++) {
        if ((r / hr) & 1) continue;
        const uint32_t e = k + (r % hr) * half;
        Fr u = x[r], v = x[r + hr] * tw[(size_t)e * (n / 2 / h)];
        x[r] = u + v;
        x[r + hr] = u - v;
      }
    }
    ZK_UNROLL for (uint32_t r = 0; r < R; r++) base[(i0 + (size_t)r * half) * B + b] = x[r];
  }
}
// after the DIF inverse transform position p holds coefficient bitrev(p): multiply by
// tab[p] = n^-1 * inc^bitrev(p)  (inc = w_{2n}: the odd-coset shift snarkjs applies with batchApplyKey)
ZK_GLOBAL void k_scale_rows(Fr* __restrict__ data, const Fr* __restrict__ tab, uint32_t n, uint32_t B, uint32_t n_poly) {
  size_t tid = ZK_TID;
  size_t per_poly = (size_t)n * B;
  if (tid >= per_poly * n_poly) return;
  uint32_t p = (uint32_t)((tid % per_poly) / B);
  data[tid] = data[tid] * tab[p];
}
// joinABC: P = A'*B' - C', Montgomery -> canonical (the H-MSM scalars), out [n][B]
ZK_GLOBAL void k_join_abc(const Fr* __restrict__ abc, Fr* __restrict__ out, uint32_t n, uint32_t B) {
  size_t tid = ZK_TID;
  size_t per_poly = (size_t)n * B;
  if (tid >= per_poly) return;
  Fr a = abc[tid], b = abc[per_poly + tid], c = abc[2 * per_poly + tid];
  out[tid] = (a * b - c).from_mont();
}

#endif  // ZK_K_WITNESS
// ================================================================================ K6/K7: Pippenger MSM
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
// skey: the bucket index of every entry (lets pass 4 walk the list in fixed-size chunks)
ZK_GLOBAL void k_msm_scatter(const Fr* __restrict__ scalars, const uint8_t* __restrict__ skip, MsmShape s,
                             uint32_t* __restrict__ cursors, uint32_t* __restrict__ sorted, uint16_t* __restrict__ skey) {
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
    skey[at] = (uint16_t)(mag - 1);
  }
}
#endif  // ZK_K_MSM_SORT
// pass 4: BALANCED bucket accumulation. One thread per (row, chunk of S consecutive sorted entries): every thread
// performs exactly S mixed adds whatever the bucket-size distribution (witness scalars are far from uniform:
// bits, small values, and the partial top window concentrate thousands of entries in a few buckets).
// A bucket lying inside one chunk is written directly; a bucket cut by chunk borders leaves partial sums
// (head = run containing the chunk's first entry, tail = run containing its last entry) for pass 4b.
template <class F>
ZK_GLOBAL ZK_ACC_BOUNDS(F) void k_msm_accumulate_chunks(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                       const uint16_t* __restrict__ skey, const uint32_t* __restrict__ offsets,
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
  const uint16_t* keys = skey + row * s.cap;
  uint32_t cur = keys[msm_list_index(s, pos0)];
  bool first = true;
  Xyzz<F> acc = Xyzz<F>::infinity();
  // one entry of lookahead: the next key / list word are loaded, and the next base prefetched, while this entry's mixed add runs
  uint32_t k_next = cur, e_next = list[msm_list_index(s, pos0)], e_next2 = 0;
  if (pos0 + 1 < pos1) e_next2 = list[msm_list_index(s, pos0 + 1)];
  for (uint32_t pos = pos0; pos < pos1; pos++) {
    const uint32_t k = k_next, e = e_next;
    e_next = e_next2;
    if (pos + 1 < pos1) {
      ZK_PREFETCH(bases + (e_next & 0x7FFFFFFFu));
      k_next = keys[msm_list_index(s, pos + 1)];
      if (pos + 2 < pos1) e_next2 = list[msm_list_index(s, pos + 2)];
    }
    if (k != cur) {
      // the run of bucket `cur` ends inside this chunk; it is whole unless it began in an earlier chunk
      if (first && off[cur] < pos0) head[tid] = acc; else buckets[row * s.nb + cur] = acc;
      acc = Xyzz<F>::infinity();
      cur = k;
      first = false;
    }
    xyzz_madd(acc, bases[e & 0x7FFFFFFFu], (e >> 31) != 0);
  }
  bool starts_here = !(first && off[cur] < pos0);
  bool ends_here = off[cur] + cnt[cur] <= pos1;
  if (starts_here && ends_here) buckets[row * s.nb + cur] = acc;
  else if (first) head[tid] = acc;   // run covers the chunk's first entry (possibly the whole chunk)
  else tail[tid] = acc;              // run started here and continues in the next chunk
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
                                       const uint16_t* __restrict__ skey, const uint32_t* __restrict__ offsets,
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
  const uint16_t* row_keys = skey + (size_t)row * s.cap;
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
template <class F>
ZK_GLOBAL void k_msm_fixup(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts, MsmShape s, uint32_t S,
                           uint32_t chunks_per_row, const Xyzz<F>* __restrict__ head, const Xyzz<F>* __restrict__ tail,
                           Xyzz<F>* __restrict__ buckets) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)s.B * s.R * s.nb) return;
  size_t row = tid / s.nb;
  uint32_t st = offsets[tid], cnt = counts[tid];
  if (cnt == 0) { buckets[tid] = Xyzz<F>::infinity(); return; }
  uint32_t c0 = st / S, c1 = (st + cnt - 1) / S;
  if (c0 == c1) return;  // written whole by its chunk
  const Xyzz<F>* h = head + row * chunks_per_row;
  const Xyzz<F>* t = tail + row * chunks_per_row;
  Xyzz<F> acc = (st > c0 * S) ? t[c0] : h[c0];
  for (uint32_t ch = c0 + 1; ch <= c1; ch++) xyzz_add(acc, h[ch]);
  buckets[tid] = acc;
}
// pass 5: bucket reduction S = sum_k (k+1) * X[k] over the nb buckets of a row, as a three-level tree so that the
// serial depth is ~ 2*L1 + 2*L2 + 5*N2 additions instead of 2*sqrt(nb) + 3*sqrt(nb).
// Level kernel: chunk t of L consecutive elements -> R_t = sum X, T_t = sum j * X[t*L + j] (zero-based local weights).
// With Z(X) = sum_k k * X[k]:  Z(X) = sum_t T_t + L * Z(R)  and  S = Z(X) + sum(X) = Z(X) + sum(R).
template <class F>
ZK_GLOBAL ZK_RED_BOUNDS void k_reduce_level(const Xyzz<F>* __restrict__ in, size_t rows, uint32_t N, uint32_t L, Xyzz<F>* __restrict__ R,
                              Xyzz<F>* __restrict__ T) {
  size_t tid = ZK_TID;
  uint32_t nchunk = N / L;
  if (tid >= rows * nchunk) return;
  size_t row = tid / nchunk;
  uint32_t ch = (uint32_t)(tid % nchunk);
  const Xyzz<F>* x = in + row * N + (size_t)ch * L;
  Xyzz<F> run = Xyzz<F>::infinity(), acc = Xyzz<F>::infinity();
  ZK_NOUNROLL for (int j = (int)L - 1; j >= 1; j--) {
    xyzz_add_hot(run, x[j]);
    xyzz_add_hot(acc, run);
  }
  xyzz_add_hot(run, x[0]);
  R[tid] = run;
  if (T) T[tid] = acc;
}
// Level 1 of the tree with the FIX-UP FUSED IN (batch path): the value of bucket k is taken straight from what the accumulation
// left -- nothing for an empty bucket, buckets[k] when one chunk held the whole run, otherwise the partial sums of the chunks
// the run crosses (tail of the first, heads of the following ones).  Every partial is added to the running sum directly, so a
// bucket cut by chunk borders costs the same additions as before but no separate pass over all buckets, and the bucket array is
// neither completed nor re-read.  Thread = (row, chunk of L buckets).
template <class F>
ZK_GLOBAL ZK_RED_BOUNDS void k_reduce_level1_fused(const Xyzz<F>* __restrict__ buckets, const Xyzz<F>* __restrict__ head,
                                                   const Xyzz<F>* __restrict__ tail, const uint32_t* __restrict__ offsets,
                                                   const uint32_t* __restrict__ counts, size_t rows, uint32_t nb, uint32_t L, uint32_t S,
                                                   uint32_t chunks_per_row, Xyzz<F>* __restrict__ R, Xyzz<F>* __restrict__ T) {
  size_t tid = ZK_TID;
  const uint32_t nchunk = nb / L;
  if (tid >= rows * nchunk) return;
  const size_t row = tid / nchunk;
  const uint32_t ch = (uint32_t)(tid % nchunk);
  const uint32_t* off = offsets + row * nb + (size_t)ch * L;
  const uint32_t* cnt = counts + row * nb + (size_t)ch * L;
  const Xyzz<F>* x = buckets + row * nb + (size_t)ch * L;
  const Xyzz<F>* h = head + row * chunks_per_row;
  const Xyzz<F>* t = tail + row * chunks_per_row;
  Xyzz<F> run = Xyzz<F>::infinity(), acc = Xyzz<F>::infinity();
  ZK_NOUNROLL for (int j = (int)L - 1; j >= 0; j--) {
    const uint32_t n = ZK_LDG(cnt + j);
    if (n) {
      const uint32_t st = ZK_LDG(off + j), c0 = st / S, c1 = (st + n - 1) / S;
      if (c0 == c1) xyzz_add_hot(run, x[j]);
      else {
        xyzz_add_hot(run, (st > c0 * S) ? t[c0] : h[c0]);
        ZK_NOUNROLL for (uint32_t q = c0 + 1; q <= c1; q++) xyzz_add_hot(run, h[q]);
      }
    }
    if (j >= 1) xyzz_add_hot(acc, run);
  }
  R[tid] = run;
  T[tid] = acc;
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
    if (bit == 0xFFFFFFFFu || ((j >> bit) & 1u)) xyzz_add(acc, src[j]);
  *dst = acc;
}
// main: [rows] (Y_all), pool: [n_bits][rows] (Y_b): out[row] = Y_all + sum_b 2^b Y_b by Horner from the top bit
template <class F>
ZK_GLOBAL void k_reduce_bits_final(const Xyzz<F>* __restrict__ main_in, const Xyzz<F>* __restrict__ pool, uint32_t n_bits, size_t rows,
                                   Xyzz<F>* __restrict__ out) {
  size_t row = ZK_TID;
  if (row >= rows) return;
  Xyzz<F> z = Xyzz<F>::infinity();
  for (int b = (int)n_bits - 1; b >= 0; b--) {
    z = xyzz_dbl(z);
    xyzz_add(z, pool[(size_t)b * rows + row]);
  }
  xyzz_add(z, main_in[row]);
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

// ================================================================================ K8: blinding / finalisation
struct VkDev {
  G1Affine alpha1, beta1, delta1;
  G2Affine beta2, delta2;
};
#ifdef ZK_K_FIN
// phase 1: fixed-base terms. thread (b, k): k=0 r*delta1, 1 s*delta1, 2 -(r*s)*delta1 (G1) ; k=3 s*delta2 (G2)
// rs: canonical [B][2] (host order: r then s). t_g1: [B][3], t_g2: [B]
ZK_GLOBAL void k_fin_fixed(const G1Affine* __restrict__ tab_d1, const G2Affine* __restrict__ tab_d2, const Fr* __restrict__ rs,
                           uint32_t B, G1Xyzz* __restrict__ t_g1, G2Xyzz* __restrict__ t_g2) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)B * 4) return;
  uint32_t b = (uint32_t)(tid / 4), k = (uint32_t)(tid % 4);
  Fr r = rs[2 * (size_t)b], sv = rs[2 * (size_t)b + 1];
  if (k == 0) t_g1[3 * (size_t)b] = fixed_base_mul<Fq>(tab_d1, r.v);
  else if (k == 1) t_g1[3 * (size_t)b + 1] = fixed_base_mul<Fq>(tab_d1, sv.v);
  else if (k == 2) {
    Fr nrs = (r.to_mont() * sv.to_mont()).neg().from_mont();
    t_g1[3 * (size_t)b + 2] = fixed_base_mul<Fq>(tab_d1, nrs.v);
  } else t_g2[b] = fixed_base_mul<Fq2>(tab_d2, sv.v);
}
// phase 2: thread (b, k): k=0: pi_a = A + alpha + r*delta1, then s*pi_a ; k=1: pi_b1 = B1 + beta1 + s*delta1, then r*pi_b1
// msm_g1: [4][B] = A, B1, C, H results. outputs pis[B][2] (pi_a, pi_b1) and var[B][2] (s*pi_a, r*pi_b1)
ZK_GLOBAL void k_fin_var(VkDev vk, const Fr* __restrict__ rs, uint32_t B, const G1Xyzz* __restrict__ msm_g1,
                         const G1Xyzz* __restrict__ t_g1, G1Xyzz* __restrict__ pis, G1Xyzz* __restrict__ var) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)B * 2) return;
  uint32_t b = (uint32_t)(tid / 2), k = (uint32_t)(tid % 2);
  Fr r = rs[2 * (size_t)b], sv = rs[2 * (size_t)b + 1];
  G1Xyzz p = msm_g1[(size_t)k * B + b];
  xyzz_madd(p, k == 0 ? vk.alpha1 : vk.beta1, false);
  xyzz_add(p, t_g1[3 * (size_t)b + k]);
  pis[tid] = p;
  var[tid] = xyzz_scalar_mul(p, k == 0 ? sv.v : r.v);
}
// phase 3: assemble, normalise, write the 256-byte proof (A | B | C, affine canonical LE)
ZK_GLOBAL void k_fin_write(VkDev vk, uint32_t B, const G1Xyzz* __restrict__ msm_g1, const G2Xyzz* __restrict__ msm_g2,
                           const G1Xyzz* __restrict__ t_g1, const G2Xyzz* __restrict__ t_g2, const G1Xyzz* __restrict__ pis,
                           const G1Xyzz* __restrict__ var, Fq* __restrict__ proofs /* [B][8] */) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)B * 3) return;
  uint32_t b = (uint32_t)(tid / 3), k = (uint32_t)(tid % 3);
  Fq* out = proofs + 8 * (size_t)b;
  if (k == 0) {
    G1Affine a = xyzz_to_affine(pis[2 * (size_t)b]);
    out[0] = a.x.from_mont(); out[1] = a.y.from_mont();
  } else if (k == 1) {
    G2Xyzz p = msm_g2[b];
    xyzz_madd(p, vk.beta2, false);
    xyzz_add(p, t_g2[b]);
    G2Affine a = xyzz_to_affine(p);
    out[2] = a.x.a.from_mont(); out[3] = a.x.b.from_mont(); out[4] = a.y.a.from_mont(); out[5] = a.y.b.from_mont();
  } else {
    G1Xyzz p = msm_g1[2 * (size_t)B + b];
    xyzz_add(p, msm_g1[3 * (size_t)B + b]);
    xyzz_add(p, var[2 * (size_t)b]);
    xyzz_add(p, var[2 * (size_t)b + 1]);
    xyzz_add(p, t_g1[3 * (size_t)b + 2]);
    G1Affine a = xyzz_to_affine(p);
    out[6] = a.x.from_mont(); out[7] = a.y.from_mont();
  }
}

#endif  // ZK_K_FIN
// ================================================================================ V1: batch Groth16 verifier
#ifdef ZK_K_VERIFY
// SURVEY 8f item 1 (Server.verify*Proof, tests/full_system_simulation.mjs:848-1131: one `snarkjs groth16 verify` process per
// proof).  B proofs under one verification key; the work of a proof is split over threads so that a whole round's proofs run
// concurrently: (b, j) public-input scalar multiplications, (b) decoding / curve checks / vk_x, (b, pair) Miller loops,
// (b, side) the two 761-bit halves of the final exponentiation (pairing.cuh), (b) comparison.  Latency-bound (each thread is
// a serial chain of ~25 k / ~140 k Montgomery products): throughput comes from B, not from the single proof.
ZK_GLOBAL void k_vfy_ic_mul(const G1Affine* __restrict__ ic, const Fr* __restrict__ publics, uint32_t l, uint32_t B,
                            G1Xyzz* __restrict__ t) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)B * l) return;
  const uint32_t j = (uint32_t)(tid % l);
  const Fr s = publics[tid];                        // host layout [b][j], canonical
  t[tid] = xyzz_scalar_mul(G1Xyzz::from_affine(ic[j + 1]), s.v);
}
// g1s: [B][3] = (-A, vk_x, C); g2b: [B] = B; flags[b] = 0 when the proof is malformed (coordinate >= q, public >= r, off-curve)
ZK_GLOBAL void k_vfy_prepare(zkp::PairingConsts k, const G1Affine* __restrict__ ic, const Fr* __restrict__ publics, uint32_t l,
                             uint32_t B, const uint32_t* __restrict__ proofs, const G1Xyzz* __restrict__ t,
                             zkp::G1P* __restrict__ g1s, zkp::G2P* __restrict__ g2b, uint32_t* __restrict__ flags) {
  size_t b = ZK_TID;
  if (b >= B) return;
  // everything is decoded straight from global memory: a first version staged the proof words and each public signal in
  // local arrays and came back with A.x = 0 on the sm_100a build only (the host emulation and tests/dev/neg_probe.cu, which
  // isolates the pieces, were correct) -- no local staging arrays here.
  const uint32_t* pw = proofs + b * 64;
  bool ok = true;
  ZK_NOUNROLL for (int i = 0; i < 8; i++) ok = ok && zkp::canonical_lt(pw + 8 * i, false);
  ZK_NOUNROLL for (uint32_t j = 0; j < l; j++) ok = ok && zkp::canonical_lt(publics[b * l + j].v, true);
  zkp::G1P A = zkp::g1_from_canonical(pw), C = zkp::g1_from_canonical(pw + 48);
  zkp::G2P Bp = zkp::g2_from_canonical(pw + 16);
  ok = ok && !A.inf && !C.inf && !Bp.inf;          // (0, 0) is not on the curve: malformed, as in the host verifier
  ok = ok && zkp::g1_on_curve(A, k) && zkp::g1_on_curve(C, k) && zkp::g2_on_curve(Bp, k);
  A.y = A.y.neg();
  g1s[3 * b] = A;
  g1s[3 * b + 2] = C;
  g2b[b] = Bp;
  flags[b] = ok ? 1u : 0u;
  G1Xyzz vkx = G1Xyzz::from_affine(ic[0]);
  ZK_NOUNROLL for (uint32_t j = 0; j < l; j++) xyzz_add(vkx, t[b * l + j]);
  g1s[3 * b + 1] = zkp::g1_from_xyzz(vkx);
}
// thread t < 3B: Miller loop of pair t % 3 of proof t / 3 ((B, -A), (gamma, vk_x), (delta, C)); thread 3B: (beta, alpha), shared
ZK_GLOBAL void k_vfy_miller(zkp::PairingConsts k, zkp::G2P beta, zkp::G2P gamma, zkp::G2P delta, zkp::G1P alpha,
                            const zkp::G1P* __restrict__ g1s, const zkp::G2P* __restrict__ g2b, uint32_t B,
                            zkp::F12* __restrict__ f, uint32_t* __restrict__ flags, int affine) {
  size_t tid = ZK_TID;
  if (tid > (size_t)3 * B) return;
  const size_t b = tid < (size_t)3 * B ? tid / 3 : 0;
  const uint32_t pair = tid < (size_t)3 * B ? (uint32_t)(tid % 3) : 3u;
  if (pair < 3 && !flags[b]) return;                // malformed input: nothing to pair
  // one code path for all four kinds of pair (selected operands, no divergent copies of the loop)
  const zkp::G2P Q = pair == 0 ? g2b[b] : pair == 1 ? gamma : pair == 2 ? delta : beta;
  const zkp::G1P P = pair < 3 ? g1s[tid] : alpha;
  zkp::F12 out;
  if (affine) {                                     // cross-check form: affine line steps in the flat basis
    if (!zkp::miller(Q, P, out, k)) { if (pair < 3) flags[b] = 0; return; }
  } else {
    zkp::miller_proj(Q, P, out, k);
  }
  f[tid] = out;
}
// thread (b, side): one half of the final exponentiation of F_b = f[3b] f[3b+1] f[3b+2] f[3B]
ZK_GLOBAL void k_vfy_final(zkp::PairingConsts k, const zkp::F12* __restrict__ f, const uint32_t* __restrict__ flags, uint32_t B,
                           zkp::F12* __restrict__ halves) {
  // a warp works on ONE side (32 proofs): lanes never diverge on it
  const size_t tid = ZK_TID;
  const size_t b = (tid >> 6) * 32 + (tid & 31);
  const int side = (int)((tid >> 5) & 1);
  if (b >= B) return;
  if (!flags[b]) return;
  zkp::F12 F = f[3 * b], g;
  ZK_NOUNROLL for (int i = 1; i < 4; i++) { g = f[i < 3 ? 3 * b + i : (size_t)3 * B]; zkp::f12_mul(F, F, g, k); }
  zkp::final_half(F, side, g, k);
  halves[2 * b + side] = g;
}
// thread b: the whole final exponentiation of proof b in the tower view (easy part + x-power chain, ~20 k products)
ZK_GLOBAL void k_vfy_final_tower(zkp::PairingConsts k, const zkp::F12* __restrict__ f, const uint32_t* __restrict__ flags, uint32_t B,
                                 int32_t* __restrict__ ok) {
  const size_t b = ZK_TID;
  if (b >= B) return;
  if (!flags[b]) { ok[b] = 0; return; }
  zkp::F12 F = f[3 * b], g;
  ZK_NOUNROLL for (int i = 1; i < 4; i++) { g = f[i < 3 ? 3 * b + i : (size_t)3 * B]; zkp::f12_mul(F, F, g, k); }
  ok[b] = zkp::final_exp_is_one(F, k) ? 1 : 0;
}
ZK_GLOBAL void k_vfy_compare(const zkp::F12* __restrict__ halves, const uint32_t* __restrict__ flags, uint32_t B,
                             int32_t* __restrict__ ok) {
  size_t b = ZK_TID;
  if (b >= B) return;
  ok[b] = (flags[b] && zkp::f12_eq(halves[2 * b], halves[2 * b + 1])) ? 1 : 0;
}

#endif  // ZK_K_VERIFY
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
#ifdef ZK_K_BENCH
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

#endif  // ZK_K_BENCH

}  // namespace zk
