// sm_100a kernels of the key generation (`snarkjs groth16 setup`, tests/full_system_simulation.mjs:714-717):
// Lagrange basis at tau, the per-wire column sums of A, B, C against it, the key scalars, and fixed-base multiplications of
// the generators from byte-window tables with a shared inversion for the affine conversion.  Only setup.cu includes it.
#pragma once
#include "k_msm.cuh"

namespace zk {

ZK_HD Fr fr_pow_u32(Fr base, uint32_t e) {
  Fr r = Fr::one();
  ZK_NOUNROLL for (int i = 31; i >= 0; i--) { r = r.sqr(); if ((e >> i) & 1u) r = r * base; }
  return r;
}

// out[k] = scale * e_k / (tau - e_k), e_k = first * step^k, k < n   (all Montgomery).
// Lagrange basis of the size-n domain at tau: first = 1, step = w_n, scale = (tau^n - 1) / n.
// Odd half of the size-2n domain (the H query of snarkjs): first = w_2n, step = w_n, scale = (tau^2n - 1) / 2n.
// A thread owns ZK_LAG_CH consecutive k and shares ONE inversion among them (Montgomery's trick).
#define ZK_LAG_CH 16
ZK_GLOBAL void k_lagrange(Fr tau, Fr first, Fr step, Fr scale, uint32_t n, Fr* __restrict__ out) {
  const size_t t = ZK_TID;
  const size_t k0 = t * ZK_LAG_CH;
  if (k0 >= n) return;
  const uint32_t cnt = n - k0 < ZK_LAG_CH ? (uint32_t)(n - k0) : ZK_LAG_CH;
  Fr e[ZK_LAG_CH], pre[ZK_LAG_CH];
  Fr cur = first * fr_pow_u32(step, (uint32_t)k0), prod = Fr::one();
  ZK_NOUNROLL for (uint32_t k = 0; k < cnt; k++) {
    e[k] = cur;
    pre[k] = prod;
    prod = prod * (tau - cur);
    cur = cur * step;
  }
  Fr inv = prod.inv();
  ZK_NOUNROLL for (int k = (int)cnt - 1; k >= 0; k--) {
    const Fr dinv = inv * pre[k];
    inv = inv * (tau - e[k]);
    out[k0 + k] = scale * e[k] * dinv;
  }
}

// column sums, step 1: piece p covers entries [piece_off[p], piece_off[p+1]) of the by-wire ordering (one wire per piece):
// partial[p] = sum coef[cidx[e]] * L[row[e]]
ZK_GLOBAL void k_colsum_pieces(const uint32_t* __restrict__ piece_off, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ cidx,
                               const Fr* __restrict__ coef, const Fr* __restrict__ L, size_t n_pieces, Fr* __restrict__ partial) {
  const size_t p = ZK_TID;
  if (p >= n_pieces) return;
  Fr acc = Fr::zero();
  const uint32_t hi = ZK_LDG(piece_off + p + 1);
  for (uint32_t e = ZK_LDG(piece_off + p); e < hi; e++) acc = acc + coef[ZK_LDG(cidx + e)] * L[ZK_LDG(rows + e)];
  partial[p] = acc;
}
// step 2: wire w owns pieces [wire_piece[w], wire_piece[w+1])
ZK_GLOBAL void k_colsum_wires(const uint32_t* __restrict__ wire_piece, const Fr* __restrict__ partial, uint32_t m, Fr* __restrict__ out) {
  const size_t w = ZK_TID;
  if (w >= m) return;
  Fr acc = Fr::zero();
  const uint32_t hi = ZK_LDG(wire_piece + w + 1);
  for (uint32_t p = ZK_LDG(wire_piece + w); p < hi; p++) acc = acc + partial[p];
  out[w] = acc;
}

// key scalars (canonical, ready for the byte windows).  At/Bt/Ct: column sums (Montgomery); L: Lagrange basis; Lodd: odd half of the
// double domain.  Layout of `sc` (G1): [alpha, beta, delta | IC: l+1 | A: m | B1: m | C: m-l-1 | H: n];  `sc2` (G2): [beta, 1, delta | B2: m].
ZK_GLOBAL void k_setup_scalars(const Fr* __restrict__ At, const Fr* __restrict__ Bt, const Fr* __restrict__ Ct, const Fr* __restrict__ L,
                               const Fr* __restrict__ Lodd, Fr alpha, Fr beta, Fr delta, Fr delta_inv, uint32_t m, uint32_t l, uint32_t nc,
                               uint32_t n, Fr* __restrict__ sc, Fr* __restrict__ sc2) {
  const size_t t = ZK_TID;
  const size_t o_ic = 3, o_a = o_ic + l + 1, o_b = o_a + m, o_c = o_b + m, o_h = o_c + (m - l - 1);
  if (t < 3) {
    const Fr v = t == 0 ? alpha : t == 1 ? beta : delta;
    sc[t] = v.from_mont();
    sc2[t] = (t == 0 ? beta : t == 1 ? Fr::one() : delta).from_mont();
  }
  if (t < m) {
    Fr a = At[t];
    if (t <= l) a = a + L[nc + t];                // the extra A rows that bind the public inputs (snarkjs setup)
    const Fr b = Bt[t];
    const Fr comb = beta * a + alpha * b + Ct[t];
    sc[o_a + t] = a.from_mont();
    sc[o_b + t] = b.from_mont();
    sc2[3 + t] = b.from_mont();
    if (t <= l) sc[o_ic + t] = comb.from_mont();  // gamma = 1
    else sc[o_c + (t - l - 1)] = (comb * delta_inv).from_mont();
  }
  if (t < n) sc[o_h + t] = (Lodd[t] * delta_inv).from_mont();
}

// out[i] = k_i * G from the byte-window table of G (32 mixed additions), affine Montgomery (zkey point layout).  A thread makes
// ZK_FB_CH points and converts them to affine with ONE inversion.
#define ZK_FB_CH 4
template <class F>
ZK_GLOBAL void k_fixed_base_batch(const Affine<F>* __restrict__ tab, const Fr* __restrict__ scalars, size_t n, Affine<F>* __restrict__ out) {
  const size_t i0 = ZK_TID * ZK_FB_CH;
  if (i0 >= n) return;
  const uint32_t cnt = n - i0 < ZK_FB_CH ? (uint32_t)(n - i0) : ZK_FB_CH;
  Xyzz<F> p[ZK_FB_CH];
  F pre[ZK_FB_CH];
  F prod = F::one();
  ZK_NOUNROLL for (uint32_t k = 0; k < cnt; k++) {
    const Fr s = scalars[i0 + k];
    p[k] = fixed_base_mul<F>(tab, s.v);
    pre[k] = prod;
    if (!p[k].is_inf()) prod = prod * p[k].ZZZ;
  }
  F inv = prod.inv();
  ZK_NOUNROLL for (int k = (int)cnt - 1; k >= 0; k--) {
    Affine<F> a;
    if (p[k].is_inf()) { a.x = F::zero(); a.y = F::zero(); }
    else {
      const F i3 = inv * pre[k];                   // 1 / ZZZ_k
      inv = inv * p[k].ZZZ;
      const F i2 = p[k].ZZ.sqr() * i3.sqr();       // 1 / ZZ = ZZ^2 / ZZZ^2
      a.x = p[k].X * i2;
      a.y = p[k].Y * i3;
    }
    out[i0 + k] = a;
  }
}

}  // namespace zk
