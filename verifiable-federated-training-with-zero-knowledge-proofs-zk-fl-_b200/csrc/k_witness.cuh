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
// This header: layout transposes, the batched witness evaluator (W1), A.w / B.w (K1) and the H polynomial (K2-K5); only witness.cu includes it.
#pragma once
#include "types.cuh"

namespace zk {

// ================================================================================ layout helpers
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

// ================================================================================ W1: batched witness evaluator

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

// ================================================================================ K1: sparse A.w, B.w, C = A o B
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
// K1 with the constraint check folded in (full-prove path): rows below n_rows are the circuit's constraints (the zkey's A / B
// coefficient rows ARE the R1CS rows there; rows above only bind the public inputs), so A.w and B.w are computed once and only
// C.w is extra work.  first_bad as in k_r1cs_check.
ZK_GLOBAL void k_build_abc_check(CsrDev A, CsrDev Bm, CsrDev C, const Fr* __restrict__ w, Fr* __restrict__ abc, uint32_t n, uint32_t B,
                                 uint32_t n_rows, uint32_t* __restrict__ first_bad) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)n * B) return;
  uint32_t row = (uint32_t)(tid / B), b = (uint32_t)(tid % B);
  Fr a = csr_row(A, row, w, B, b), bv = csr_row(Bm, row, w, B, b);
  const Fr ab = a * bv;
  abc[tid] = a;
  abc[(size_t)n * B + tid] = bv;
  abc[2 * (size_t)n * B + tid] = ab;
  if (row < n_rows && !(ab == csr_row(C, row, w, B, b))) ZK_ATOMIC_MIN(first_bad + b, row);
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

// ================================================================================ masked aggregation + model update
// Server.aggregateUpdates (tests/full_system_simulation.mjs:1137-1199): field sum of the accepted clients' masked updates per
// model coordinate (the pairwise masks cancel), signed decode (values above r/2 are negatives), mean, SGD step.
#define ZK_AGG_CHUNK 64u
// thread (j, chunk): partial[j * n_chunks + chunk] = sum over the chunk's accepted clients of masked[client][j]  (canonical mod-r sums);
// flag[0] |= 1 when a value is not reduced mod r
ZK_GLOBAL void k_agg_partial(const Fr* __restrict__ masked, const uint8_t* __restrict__ accept, uint32_t n_clients, uint32_t dim,
                             uint32_t n_chunks, Fr* __restrict__ partial, uint32_t* __restrict__ flag) {
  const size_t tid = ZK_TID;
  if (tid >= (size_t)dim * n_chunks) return;
  const uint32_t j = (uint32_t)(tid / n_chunks), ch = (uint32_t)(tid % n_chunks);
  const uint32_t lo = ch * ZK_AGG_CHUNK, hi = lo + ZK_AGG_CHUNK < n_clients ? lo + ZK_AGG_CHUNK : n_clients;
  Fr acc = Fr::zero();
  bool bad = false;
  for (uint32_t i = lo; i < hi; i++) {
    if (accept && !accept[i]) continue;
    const Fr v = masked[(size_t)i * dim + j];
    bool lt = false, decided = false;
    ZK_UNROLL for (int w = 7; w >= 0; w--) if (!decided && v.v[w] != FrP::mod(w)) { lt = v.v[w] < FrP::mod(w); decided = true; }
    bad = bad || !lt;
    acc = acc + v;
  }
  partial[tid] = acc;
  if (bad) ZK_ATOMIC_OR(flag, 1u);
}
// magnitude (8 little-endian words) -> double with round-to-nearest-even, what JavaScript's Number(BigInt) does
ZK_D double u256_to_double(const uint32_t* w) {
  int top = -1;
  ZK_UNROLL for (int i = 7; i >= 0; i--) if (top < 0 && w[i]) top = i;
  if (top < 0) return 0.0;
  if (top <= 1) return (double)(((uint64_t)w[1] << 32) | w[0]);           // fits 64 bits: the conversion itself rounds to nearest even
  // 64-bit window ending at the most significant word, sticky bit for everything below it
  const uint64_t hi = ((uint64_t)w[top] << 32) | w[top - 1];
  bool sticky = false;
  for (int i = 0; i < top - 1; i++) sticky = sticky || w[i] != 0;
  // normalise so that bit 63 is set, keeping the bits shifted in from the next word
  int lz = 0;
  while (!((hi << lz) >> 63)) lz++;
  uint64_t win = hi << lz;
  if (lz) {
    const uint32_t nxt = w[top - 2];
    win |= (uint64_t)nxt >> (32 - lz);
    sticky = sticky || (uint32_t)(nxt << lz) != 0;
    for (int i = 0; i < top - 2; i++) sticky = sticky || w[i] != 0;
  }
  if (sticky) win |= 1u;                                                   // 64 > 53 + 2 bits: the sticky bit keeps the rounding exact
  double d = (double)win;
  const int shift = 32 * (top - 1) - lz;                                   // value = win * 2^shift
  for (int i = 0; i < shift; i++) d *= 2.0;
  return d;
}
#if defined(__CUDA_ARCH__)
#define ZK_DMUL(a, b) __dmul_rn((a), (b))      // no fused multiply-add: the reference rounds the product, then the difference
#define ZK_DSUB(a, b) __dsub_rn((a), (b))
#else
#define ZK_DMUL(a, b) ((double)((volatile double)(a) * (b)))
#define ZK_DSUB(a, b) ((a) - (b))
#endif
// thread j: total, signed decode, mean over `count` clients, model_out[j] = model_in[j] - lr * mean
ZK_GLOBAL void k_agg_final(const Fr* __restrict__ partial, uint32_t n_chunks, uint32_t dim, uint32_t count, double lr,
                           const double* __restrict__ model_in, Fr* __restrict__ agg_field, double* __restrict__ mean,
                           double* __restrict__ model_out) {
  const size_t j = ZK_TID;
  if (j >= dim) return;
  Fr acc = Fr::zero();
  for (uint32_t c = 0; c < n_chunks; c++) acc = acc + partial[j * n_chunks + c];
  agg_field[j] = acc;
  // aggregatedMasked > FIELD_PRIME / 2n  (BigInt division: (r - 1) / 2)  <=>  2 * acc > r - 1  <=>  2 * acc >= r + 1 ... compare acc with half
  uint32_t half[8];
  ZK_UNROLL for (int i = 0; i < 8; i++) half[i] = (FrP::mod(i) >> 1) | (i < 7 ? FrP::mod(i + 1) << 31 : 0u);   // (r - 1) / 2 (r is odd)
  bool gt = false, decided = false;
  ZK_UNROLL for (int i = 7; i >= 0; i--) if (!decided && acc.v[i] != half[i]) { gt = acc.v[i] > half[i]; decided = true; }
  double g;
  if (gt) { const Fr neg = Fr::zero() - acc; g = -u256_to_double(neg.v); }   // r - acc: the magnitude of the negative value
  else g = u256_to_double(acc.v);
  g = g / (double)count;
  mean[j] = g;
  model_out[j] = ZK_DSUB(model_in[j], ZK_DMUL(lr, g));
}

}  // namespace zk
