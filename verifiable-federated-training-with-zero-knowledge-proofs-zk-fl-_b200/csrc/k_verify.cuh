// sm_100a kernels of the batch Groth16 verifier; only verify.cu includes it.
#pragma once
#include "types.cuh"
#include "pairing_coop.cuh"
#include "k_msm.cuh"   // fixed_base_mul / k_fixed_base_table

namespace zk {

// ================================================================================ V1: batch Groth16 verifier
// SURVEY 8f item 1 (Server.verify*Proof, tests/full_system_simulation.mjs:848-1131: one `snarkjs groth16 verify` process per
// proof).  B proofs under one verification key; the work of a proof is split over threads so that a whole round's proofs run
// concurrently: (b, j) public-input scalar multiplications, (b) decoding / curve checks / vk_x, (b, pair) Miller loops,
// (b, side) the two 761-bit halves of the final exponentiation (pairing.cuh), (b) comparison.  Latency-bound (each thread is
// a serial chain of ~25 k / ~140 k Montgomery products): throughput comes from B, not from the single proof.
// tabs != NULL: byte-window tables of the key's points (layout [IC_0, IC_1 .. IC_l, alpha] x 8192 entries): 32 mixed additions
ZK_GLOBAL void k_vfy_ic_mul(const G1Affine* __restrict__ ic, const Fr* __restrict__ publics, uint32_t l, uint32_t B,
                            G1Xyzz* __restrict__ t, const G1Affine* __restrict__ tabs) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)B * l) return;
  const uint32_t j = (uint32_t)(tid % l);
  const Fr s = publics[tid];                        // host layout [b][j], canonical
  bool reduced = zkp::canonical_lt(s.v, true);      // an unreduced signal (the proof is rejected anyway) must not index past the table
  if (tabs && reduced) t[tid] = fixed_base_mul(tabs + (size_t)(j + 1) * 8192, s.v);
  else t[tid] = xyzz_scalar_mul(G1Xyzz::from_affine(ic[j + 1]), s.v);
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

// ================================================================================ V1, lane-cooperative form (default)
// The kernels above are one serial chain per thread (24 ms for any batch up to ~8 k proofs).  Here (pairing_coop.cuh):
//   * the G2 side of every Miller loop runs alone, one thread per G2 point, and leaves a table of line coefficients;
//   * the Fq12 accumulator of a Miller loop and the whole final exponentiation run on GROUPS OF 8 LANES;
//   * large batches are checked under random weights (one Miller loop per proof, one final exponentiation per batch);
//     the per-proof form below gives the individual verdicts when that combined check fails.
// group index of this thread; false for the emulation's non-leading lanes (one emulated thread does a whole group)
ZK_D bool coop_group(size_t& g) {
  const size_t tid = ZK_TID;
  g = tid >> 3;
#ifdef ZKFL_EMUL
  return (tid & 7u) == 0;
#else
  return true;
#endif
}
// decode + range / curve checks of proof b (no arithmetic on the publics): g2b[b] = B, flags[b]
ZK_GLOBAL void k_vfy_check(zkp::PairingConsts k, const Fr* __restrict__ publics, uint32_t l, uint32_t B, const uint32_t* __restrict__ proofs,
                           zkp::G2P* __restrict__ g2b, uint32_t* __restrict__ flags) {
  size_t b = ZK_TID;
  if (b >= B) return;
  const uint32_t* pw = proofs + b * 64;
  bool ok = true;
  ZK_NOUNROLL for (int i = 0; i < 8; i++) ok = ok && zkp::canonical_lt(pw + 8 * i, false);
  ZK_NOUNROLL for (uint32_t j = 0; j < l; j++) ok = ok && zkp::canonical_lt(publics[b * l + j].v, true);
  const zkp::G1P A = zkp::g1_from_canonical(pw), C = zkp::g1_from_canonical(pw + 48);
  const zkp::G2P Bp = zkp::g2_from_canonical(pw + 16);
  ok = ok && !A.inf && !C.inf && !Bp.inf;
  ok = ok && zkp::g1_on_curve(A, k) && zkp::g1_on_curve(C, k) && zkp::g2_on_curve(Bp, k);
  g2b[b] = Bp;
  flags[b] = ok ? 1u : 0u;
}
// line tables: t < B: the proof's B (skipped when malformed); t = B, B+1, B+2: gamma, delta, beta of the key
ZK_GLOBAL void k_vfy_lines(zkp::PairingConsts k, const zkp::G2P* __restrict__ g2b, const uint32_t* __restrict__ flags, uint32_t B,
                           zkp::G2P gamma, zkp::G2P delta, zkp::G2P beta, zkp::LineRec* __restrict__ lines) {
  size_t t = ZK_TID;
  if (t >= (size_t)B + 3) return;
  if (t < B && !flags[t]) return;
  const zkp::G2P Q = t < B ? g2b[t] : t == B ? gamma : t == (size_t)B + 1 ? delta : beta;
  zkp::miller_lines(Q, lines + t * zkp::kMillerSteps, k);
}
// Miller accumulators on lane groups.  rlc == 0 (per proof): pair p = 3b + j uses g1s[p] and table (j == 0 ? b : B + j - 1), pair 3B is
// (alpha, beta).  rlc == 1: pair p < B + 3 uses g1s[p] and table p.  f[p] in the tower layout.
ZK_GLOBAL void k_vfy_miller_coop(const zkp::G1P* __restrict__ g1s, const zkp::LineRec* __restrict__ lines, const uint32_t* __restrict__ flags,
                                 uint32_t B, uint32_t n_pairs, int rlc, zkp::T12* __restrict__ f) {
  size_t p;
  if (!coop_group(p) || p >= n_pairs) return;
  const zkp::CoopLane L;
  size_t table = p;
  bool live = true;
  if (!rlc) {
    const size_t b = p / 3;
    const uint32_t j = (uint32_t)(p % 3);
    table = p == (size_t)3 * B ? (size_t)B + 2 : j == 0 ? b : (size_t)B + j - 1;
    live = p == (size_t)3 * B || flags[b] != 0;
  } else if (p < B) live = flags[p] != 0;
  zkp::G1P P = g1s[p];
  if (!live) P.inf = 1;
  zkp::C12 acc;
  zkp::c12_miller(acc, lines + table * zkp::kMillerSteps, P, L);
  zkp::c12_store(f + p, acc, L, true);
}
// out[g] = product of in[g * fan .. min((g + 1) * fan, n_in))
ZK_GLOBAL void k_vfy_prod_coop(const zkp::T12* __restrict__ in, uint32_t n_in, uint32_t fan, zkp::T12* __restrict__ out) {
  size_t g;
  const uint32_t n_out = (n_in + fan - 1) / fan;
  if (!coop_group(g) || g >= n_out) return;
  const zkp::CoopLane L;
  const size_t lo = g * fan, hi = lo + fan < n_in ? lo + fan : n_in;
  zkp::C12 acc, x;
  zkp::c12_load(acc, in + lo, L);
  ZK_NOUNROLL for (size_t i = lo + 1; i < hi; i++) { zkp::c12_load(x, in + i, L); zkp::c12_mul(acc, acc, x, L); }
  zkp::c12_store(out + g, acc, L, true);
}
// ok[g] = final_exp(f[g * stride] * ... * f[g * stride + cnt - 1] * f[extra]) == 1   (extra == 0xFFFFFFFF: none; flags may be NULL)
ZK_GLOBAL void k_vfy_final_coop(zkp::PairingConsts k, const zkp::T12* __restrict__ f, uint32_t stride, uint32_t cnt, uint32_t extra,
                                const uint32_t* __restrict__ flags, uint32_t n, int32_t* __restrict__ ok) {
  size_t g;
  if (!coop_group(g) || g >= n) return;
  const zkp::CoopLane L;
  if (flags && !flags[g]) { ok[g] = 0; return; }
  zkp::C12 acc, x;
  zkp::c12_load(acc, f + g * stride, L);
  ZK_NOUNROLL for (uint32_t i = 1; i < cnt; i++) { zkp::c12_load(x, f + g * stride + i, L); zkp::c12_mul(acc, acc, x, L); }
  if (extra != 0xFFFFFFFFu) { zkp::c12_load(x, f + extra, L); zkp::c12_mul(acc, acc, x, L); }
  const bool one = zkp::c12_final_exp_is_one(acc, k, L);
  ok[g] = one ? 1 : 0;
}
// per-proof form: alpha as pair 3B's G1 point (k_vfy_prepare wrote the 3B proof-dependent ones)
ZK_GLOBAL void k_vfy_put_g1(zkp::G1P p, zkp::G1P* __restrict__ dst) { if (ZK_TID == 0) *dst = p; }

// -------------------------------------------------------------------------------- random-linear-combination batch check
//   prod_b e(rho_b A_b, B_b) == e(alpha, beta)^(sum rho) * e(sum_b rho_b vk_x(b), gamma) * e(sum_b rho_b C_b, delta),
//   sum_b rho_b vk_x(b) = (sum rho) IC_0 + sum_j (sum_b rho_b x_bj) IC_j :
// one Miller loop and two 128-bit scalar multiplications per proof, l + 2 scalar multiplications, three Miller loops and ONE final
// exponentiation per batch.  rho_b: 128 random bits per proof, drawn by the host for every call; malformed proofs stay out of the sums.
// partial sums of the l + 1 weights: thread (j, chunk of 64 proofs); j < l: sum rho_b x_bj, j = l: sum rho_b
ZK_GLOBAL void k_vfy_rlc_scalars_part(const Fr* __restrict__ rho, const Fr* __restrict__ publics, const uint32_t* __restrict__ flags, uint32_t l,
                                      uint32_t B, Fr* __restrict__ part) {
  const size_t tid = ZK_TID;
  const uint32_t nch = (B + 63) / 64;
  if (tid >= (size_t)(l + 1) * nch) return;
  const uint32_t j = (uint32_t)(tid / nch), ch = (uint32_t)(tid % nch);
  Fr acc = Fr::zero();
  const uint32_t hi = (ch + 1) * 64 < B ? (ch + 1) * 64 : B;
  ZK_NOUNROLL for (uint32_t b = ch * 64; b < hi; b++) {
    if (!flags[b]) continue;
    const Fr r = rho[b];
    acc = acc + (j < l ? r.to_mont() * publics[(size_t)b * l + j] : r);   // Montgomery(rho) * canonical = canonical product
  }
  part[tid] = acc;
}
ZK_GLOBAL void k_vfy_rlc_scalars_sum(const Fr* __restrict__ part, uint32_t l, uint32_t nch, Fr* __restrict__ s) {
  const size_t j = ZK_TID;
  if (j > l) return;
  Fr acc = Fr::zero();
  ZK_NOUNROLL for (uint32_t ch = 0; ch < nch; ch++) acc = acc + part[j * nch + ch];
  s[j] = acc;
}
template <class F> ZK_HD Xyzz<F> xyzz_scalar_mul_bits(const Xyzz<F>& p, const uint32_t* k, int nbits) {
  Xyzz<F> r = Xyzz<F>::infinity();
  ZK_NOUNROLL for (int i = nbits - 1; i >= 0; i--) {
    r = xyzz_dbl(r);
    if ((k[i >> 5] >> (i & 31)) & 1) xyzz_add(r, p);
  }
  return r;
}
// everything that needs only the inputs, in ONE launch (the chains run side by side, the launch lasts as long as the longest):
// warps are assigned a role, 32 items per warp.  role 0: line tables of the proofs' B; 1: P_b = -rho_b A_b, affine -> g1s[b];
// 2: rho_b C_b -> cps[b]; 3: subgroup test of B_b -> sub[b]; 4: s_j IC_j and (sum rho) alpha -> tmul[0 .. l + 1]
// (the line tables of gamma, delta, beta and the fixed-base tables behind role 4 are per-key data: zkfl_ctx::vk_cache)
ZK_GLOBAL void k_vfy_rlc_stage1(zkp::PairingConsts k, const uint32_t* __restrict__ proofs, const Fr* __restrict__ rho,
                                const uint32_t* __restrict__ flags, uint32_t B, uint32_t l, const zkp::G2P* __restrict__ g2b,
                                const G1Affine* __restrict__ tabs, const Fr* __restrict__ s, zkp::LineRec* __restrict__ lines,
                                zkp::G1P* __restrict__ g1s, G1Xyzz* __restrict__ cps, uint32_t* __restrict__ sub, G1Xyzz* __restrict__ tmul) {
  const size_t tid = ZK_TID;
  const size_t wB = ((size_t)B + 31) / 32, wL = wB;
  const size_t warp = tid >> 5, lane = tid & 31;
  if (warp < wL) {
    const size_t t = warp * 32 + lane;
    if (t >= B || !flags[t]) return;
    zkp::miller_lines(g2b[t], lines + t * zkp::kMillerSteps, k);
  } else if (warp < wL + 2 * wB) {
    const bool isC = warp >= wL + wB;
    const size_t b = (warp - wL - (isC ? wB : 0)) * 32 + lane;
    if (b >= B) return;
    const uint32_t* pw = proofs + b * 64 + (isC ? 48 : 0);
    zkp::G1P P = zkp::g1_from_canonical(pw);
    G1Xyzz r = G1Xyzz::infinity();
    if (flags[b]) r = xyzz_scalar_mul_bits(G1Xyzz::from_affine(zkp::g1_to_affine(P)), rho[b].v, 128);
    if (isC) { cps[b] = r; return; }
    P = zkp::g1_from_xyzz(r);
    P.y = P.y.neg();
    g1s[b] = P;
  } else if (warp < wL + 3 * wB) {
    const size_t b = (warp - wL - 2 * wB) * 32 + lane;
    if (b >= B) return;
    sub[b] = (flags[b] && !zkp::g2_in_subgroup(g2b[b], k)) ? 0u : 1u;
  } else {
    const size_t j = (warp - wL - 3 * wB) * 32 + lane;
    if (j > (size_t)l + 1) return;
    // j < l: s_j IC_(j+1); j = l: (sum rho) IC_0; j = l + 1: (sum rho) alpha -- from the key's byte-window tables
    const size_t table = j < l ? j + 1 : j == l ? 0 : (size_t)l + 1;
    tmul[j] = fixed_base_mul(tabs + table * 8192, s[j < l ? j : l].v);
  }
}
// out[t] = sum of in[t * fan .. ) (one level of the tree over the rho_b C_b)
ZK_GLOBAL void k_vfy_sum_g1(const G1Xyzz* __restrict__ in, uint32_t n_in, uint32_t fan, G1Xyzz* __restrict__ out) {
  const size_t t = ZK_TID;
  const uint32_t n_out = (n_in + fan - 1) / fan;
  if (t >= n_out) return;
  const size_t lo = t * fan, hi = lo + fan < n_in ? lo + fan : n_in;
  G1Xyzz acc = in[lo];
  ZK_NOUNROLL for (size_t i = lo + 1; i < hi; i++) xyzz_add(acc, in[i]);
  out[t] = acc;
}
// the three batch-wide G1 points: g1s[B] = sum_j tmul[j] (j <= l), g1s[B + 1] = csum, g1s[B + 2] = tmul[l + 1]; bad[0] |= any sub[b] == 0
ZK_GLOBAL void k_vfy_rlc_points(const G1Xyzz* __restrict__ tmul, uint32_t l, const G1Xyzz* __restrict__ csum, uint32_t B,
                                zkp::G1P* __restrict__ g1s) {
  const size_t t = ZK_TID;
  if (t >= 3) return;
  G1Xyzz acc;
  if (t == 0) { acc = tmul[0]; ZK_NOUNROLL for (uint32_t j = 1; j <= l; j++) xyzz_add(acc, tmul[j]); }
  else if (t == 1) acc = csum[0];
  else acc = tmul[l + 1];
  g1s[(size_t)B + t] = zkp::g1_from_xyzz(acc);
}
// verdicts of a batch that passed: the well-formed proofs; all_sub[0] = 1 when every well-formed B passed the subgroup test
ZK_GLOBAL void k_vfy_rlc_verdicts(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ sub, uint32_t B, int32_t* __restrict__ ok,
                                  uint32_t* __restrict__ bad_sub) {
  const size_t b = ZK_TID;
  if (b >= B) return;
  ok[b] = flags[b] ? 1 : 0;
  if (!sub[b]) ZK_ATOMIC_OR(bad_sub, 1u);
}

}  // namespace zk
