// Host-side single-proof Groth16 verifier for the drop-in boundary (`snarkjs groth16 verify`, reference call sites
// tests/full_system_simulation.mjs:865-868,975-978,1116-1119).  It runs the same pairing code as the device batch verifier
// (pairing.cuh) on one host thread: one proof is O(1) work and latency-bound, so a single `verify` call stays on the CPU
// (SURVEY 2.4 row V1); batches go through zkfl_groth16_verify_batch on the GPU.
#pragma once
#include "pairing.cuh"
#include <cstdlib>

namespace zkv {
using namespace zkp;

static const PairingConsts& consts() { static const PairingConsts k = make_consts(); return k; }

// ONE validation of a verification key for the host verifier and the device batch verifier: every coordinate reduced mod q,
// every point (alpha, beta, gamma, delta, IC[0..l]) on its curve.  false -> both paths answer ZKFL_ERR_FORMAT.
static bool vkey_well_formed(const uint8_t alpha1[64], const uint8_t beta2[128], const uint8_t gamma2[128], const uint8_t delta2[128],
                             const uint8_t* ic, uint32_t l) {
  const PairingConsts& k = consts();
  uint32_t w[32];
  auto coords_ok = [&](const uint8_t* p, int n) { memcpy(w, p, 32 * (size_t)n); for (int i = 0; i < n; i++) if (!canonical_lt(w + 8 * i, false)) return false; return true; };
  if (!coords_ok(alpha1, 2) || !g1_on_curve(g1_from_canonical(w), k)) return false;
  const uint8_t* g2s[3] = {beta2, gamma2, delta2};
  for (int i = 0; i < 3; i++) if (!coords_ok(g2s[i], 4) || !g2_on_curve(g2_from_canonical(w), k)) return false;
  for (uint32_t i = 0; i <= l; i++) if (!coords_ok(ic + 64 * (size_t)i, 2) || !g1_on_curve(g1_from_canonical(w), k)) return false;
  return true;
}

// returns 1 = valid, 0 = invalid, negative = malformed input.  All points affine canonical little-endian.
static int groth16_verify(const uint8_t alpha1[64], const uint8_t beta2[128], const uint8_t gamma2[128], const uint8_t delta2[128],
                          const uint8_t* ic /* (l+1) x 64 */, const uint8_t* publics /* l x 32 */, uint32_t l, const uint8_t proof[256]) {
  const PairingConsts& k = consts();
  if (!vkey_well_formed(alpha1, beta2, gamma2, delta2, ic, l)) return -1;
  uint32_t pw[64], w[32];
  memcpy(pw, proof, 256);
  for (int i = 0; i < 8; i++) if (!canonical_lt(pw + 8 * i, false)) return 0;
  G1P A = g1_from_canonical(pw), C = g1_from_canonical(pw + 48);
  G2P Bp = g2_from_canonical(pw + 16);
  // an all-zero proof point is the affine point (0, 0), which is not on the curve (snarkjs rejects it): malformed, not infinity
  if (A.inf || C.inf || Bp.inf) return 0;
  if (!g1_on_curve(A, k) || !g1_on_curve(C, k) || !g2_on_curve(Bp, k)) return 0;
  memcpy(w, ic, 64);
  zk::G1Xyzz vkx = zk::G1Xyzz::from_affine(g1_to_affine(g1_from_canonical(w)));
  for (uint32_t i = 0; i < l; i++) {
    uint32_t s[8]; memcpy(s, publics + 32 * (size_t)i, 32);
    if (!canonical_lt(s, true)) return 0;   // snarkjs: public signals must be < r
    memcpy(w, ic + 64 * (size_t)(i + 1), 64);
    zk::xyzz_add(vkx, zk::xyzz_scalar_mul(zk::G1Xyzz::from_affine(g1_to_affine(g1_from_canonical(w))), s));
  }
  A.y = A.y.neg();
  memcpy(w, alpha1, 64);
  const G1P g1s[4] = {A, g1_from_canonical(w), g1_from_xyzz(vkx), C};
  G2P g2s[4]; g2s[0] = Bp;
  memcpy(w, beta2, 128); g2s[1] = g2_from_canonical(w);
  memcpy(w, gamma2, 128); g2s[2] = g2_from_canonical(w);
  memcpy(w, delta2, 128); g2s[3] = g2_from_canonical(w);
  F12 f, t, lhs, rhs;
  f12_set_one(f);
  const char* flat_env = getenv("ZKFL_VERIFY_FLAT");   // cross-check knob: affine line steps + the two-power final check
  const bool flat = flat_env && *flat_env && *flat_env != '0';
  for (int i = 0; i < 4; i++) {
    if (flat) { if (!miller(g2s[i], g1s[i], t, k)) return 0; }
    else miller_proj(g2s[i], g1s[i], t, k);
    f12_mul(f, f, t, k);
  }
  if (!flat) return final_exp_is_one(f, k) ? 1 : 0;
  final_half(f, 0, lhs, k);
  final_half(f, 1, rhs, k);
  return f12_eq(lhs, rhs) ? 1 : 0;
}

// consistency of the tower view with the flat basis on pseudo-random elements; 0 = all good, else the failing check
static int pairing_selftest() {
  const PairingConsts& k = consts();
  F12 a, b, c, d;
  Fq seed = fq_small(0x1234567u);
  for (int i = 0; i < 12; i++) { seed = seed * seed + fq_small(i + 3); a.c[i] = seed; seed = seed * seed + fq_small(77); b.c[i] = seed; }
  T12 ta, tb, tc, td;
  t12_from_flat(ta, a, k); t12_from_flat(tb, b, k);
  t12_to_flat(c, ta, k);
  if (!f12_eq(c, a)) return 1;                                   // round trip
  f12_mul(c, a, b, k); t12_mul(tc, ta, tb); t12_to_flat(d, tc, k);
  if (!f12_eq(c, d)) return 2;                                   // product
  f12_sqr(c, a, k); t12_sqr(tc, ta); t12_to_flat(d, tc, k);
  if (!f12_eq(c, d)) return 3;                                   // square
  t12_inv(tc, ta); t12_mul(tc, tc, ta);
  if (!t12_is_one(tc)) return 4;                                 // inverse
  f12_frob2(c, a, k); t12_frob(tc, ta, 2, k); t12_to_flat(d, tc, k);
  if (!f12_eq(c, d)) return 5;                                   // Frobenius^2 against the flat map
  t12_frob(tc, ta, 1, k); t12_frob(tc, tc, 1, k); t12_to_flat(d, tc, k);
  if (!f12_eq(c, d)) return 6;                                   // Frobenius o Frobenius
  t12_frob(tc, ta, 1, k); t12_frob(tc, tc, 2, k); t12_frob(td, ta, 3, k);
  t12_to_flat(c, tc, k); t12_to_flat(d, td, k);
  if (!f12_eq(c, d)) return 7;                                   // Frobenius^3
  t12_frob(tc, td, 3, k); t12_conj(td, ta);                      // Frobenius^6 = conj
  t12_to_flat(c, tc, k); t12_to_flat(d, td, k);
  if (!f12_eq(c, d)) return 8;
  t12_frob(tc, ta, 1, k); t12_frob(td, tb, 1, k); t12_mul(tc, tc, td);   // Frobenius is multiplicative
  t12_mul(td, ta, tb); t12_frob(td, td, 1, k);
  t12_to_flat(c, tc, k); t12_to_flat(d, td, k);
  if (!f12_eq(c, d)) return 9;
  // the inversion-free Miller loop against the affine one: different Miller values, same pairing value
  {
    static const uint32_t G2X0[8] = {0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu};
    static const uint32_t G2X1[8] = {0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u};
    static const uint32_t G2Y0[8] = {0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u};
    static const uint32_t G2Y1[8] = {0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u};
    uint32_t w[32];
    memcpy(w, G2X0, 32); memcpy(w + 8, G2X1, 32); memcpy(w + 16, G2Y0, 32); memcpy(w + 24, G2Y1, 32);
    const G2P Q = g2_from_canonical(w);
    uint32_t g[16] = {1, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0};
    G1P P = g1_from_canonical(g);
    if (!g2_on_curve(Q, k) || !g1_on_curve(P, k)) return 10;
    for (int rep = 0; rep < 2; rep++) {
      F12 ma, mp;
      if (!miller(Q, P, ma, k)) return 11;
      miller_proj(Q, P, mp, k);
      T12 va, vp;
      final_exp_value(ma, va, k); final_exp_value(mp, vp, k);
      t12_to_flat(c, va, k); t12_to_flat(d, vp, k);
      if (!f12_eq(c, d)) return 12 + rep;
      if (t12_is_one(va)) return 14;                             // e(G1, G2) != 1
      zk::G1Xyzz P5 = zk::xyzz_dbl(zk::xyzz_dbl(zk::G1Xyzz::from_affine(g1_to_affine(P))));   // another G1 point: 4 G
      P = g1_from_xyzz(P5);
    }
  }
  return 0;
}
}  // namespace zkv
