// sm_100a kernels of the batch Groth16 verifier; only verify.cu includes it.
#pragma once
#include "types.cuh"
#include "pairing.cuh"

namespace zk {

// ================================================================================ V1: batch Groth16 verifier
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

}  // namespace zk
