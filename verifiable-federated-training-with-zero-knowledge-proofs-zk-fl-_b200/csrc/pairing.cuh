// Groth16 verification arithmetic shared by the device batch verifier (kernels.cuh, section V1) and the single-proof host
// entry point (`snarkjs groth16 verify`, reference call sites tests/full_system_simulation.mjs:865-868,975-978,1116-1119;
// Server.verify*Proof, :848-1131).
//
// Optimal ate pairing on BN254 over Fq12 = Fq[w]/(w^12 - 18 w^6 + 82) (xi = w^6 = 9 + u), the flat basis of the oracle
// (the same representation the test oracle uses, so intermediate values can be compared), G2 arithmetic affine on the twist.  Every function is written once for host and device; the
// big ones are loops over small bodies (no unrolling, local arrays) so that a kernel holds ONE copy of the Fq12 product, the
// line step and the sparse product -- the device code has no calls with stack frames, only the Fq product leaf call.
//   * lines are sparse: l = c0 + c1 w + c3 w^3 + c7 w^7 + c9 w^9 -> 60 products instead of 144;
//   * squarings use the symmetric half: 78 products;
//   * final exponentiation: f^((p^12-1)/r) == 1  <=>  (conj(f)/f)^((p^2+1) h) == 1, h = (p^4 - p^2 + 1)/r,
//     <=>  (frob2(conj f) conj f)^h == (frob2(f) f)^h  -- no Fq12 inversion, two independent 761-bit powers (the device runs
//     them in two threads) instead of one 2816-bit power.  conj = Frobenius^6 flips the odd coefficients; frob2 = Frobenius^2
//     multiplies coefficient i by zeta^i, zeta = xi^((p^2-1)/6) in Fq (both identities are checked in tests/test_oracle_pins.py).
#pragma once
#include "bn254.cuh"

namespace zkp {
using zk::Fq;
using zk::Fq2;

struct F12 { Fq c[12]; };
struct G1P { Fq x, y; uint32_t inf; };     // Montgomery affine
struct G2P { Fq2 x, y; uint32_t inf; };
struct Line { Fq c0, c1, c3, c7, c9; };

// Montgomery-form constants, built once on the host and passed to the kernels by value
struct PairingConsts {
  Fq k3, k9, k18, k82;
  Fq zeta[12];        // zeta^i
  Fq2 twist_b;        // 3 / xi
  Fq2 g12, g13;       // xi^((p-1)/3), xi^((p-1)/2): Frobenius on the twist
};

ZK_HD Fq fq_small(uint32_t v) { Fq r = Fq::zero(); r.v[0] = v; return r.to_mont(); }
ZK_HD Fq2 fq2_conj(const Fq2& x) { Fq2 r; r.a = x.a; r.b = x.b.neg(); return r; }
static inline Fq2 fq2_pow_host(Fq2 base, const uint32_t* e, int nwords) {
  Fq2 r = Fq2::one();
  for (int i = nwords * 32 - 1; i >= 0; i--) { r = r.sqr(); if ((e[i >> 5] >> (i & 31)) & 1) r = r * base; }
  return r;
}
static inline PairingConsts make_consts() {
  PairingConsts k;
  k.k3 = fq_small(3); k.k9 = fq_small(9); k.k18 = fq_small(18); k.k82 = fq_small(82);
  static const uint32_t ZETA[8] = {0x607cfd49u, 0xe4bd44e5u, 0xbb966e3du, 0xc28f069fu, 0xe0acccb0u, 0x5e6dd9e7u, 0xe131a029u, 0x30644e72u};
  Fq z; for (int i = 0; i < 8; i++) z.v[i] = ZETA[i];
  z = z.to_mont();
  k.zeta[0] = Fq::one();
  for (int i = 1; i < 12; i++) k.zeta[i] = k.zeta[i - 1] * z;
  Fq2 xi; xi.a = k.k9; xi.b = fq_small(1);
  Fq2 three; three.a = k.k3; three.b = Fq::zero();
  k.twist_b = three * xi.inv();
  static const uint32_t E3[8] = {0x4829a9c2u, 0x69602eb2u, 0xcd7b4384u, 0xdd2b2385u, 0x808072c9u, 0xe81ac1e7u, 0xa065e00du, 0x10216f7bu};
  static const uint32_t E2[8] = {0x6c3e7ea3u, 0x9e10460bu, 0xb438e546u, 0xcbc0b548u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
  k.g12 = fq2_pow_host(xi, E3, 8);
  k.g13 = fq2_pow_host(xi, E2, 8);
  return k;
}

// ------------------------------------------------------------------------------ Fq12 (flat basis)
ZK_HD void f12_set_one(F12& r) { ZK_NOUNROLL for (int i = 0; i < 12; i++) r.c[i] = Fq::zero(); r.c[0] = Fq::one(); }
ZK_HD bool f12_eq(const F12& a, const F12& b) { bool e = true; ZK_NOUNROLL for (int i = 0; i < 12; i++) e = e && (a.c[i] == b.c[i]); return e; }
// t[12..top] folded down with w^12 = 18 w^6 - 82, result in r
ZK_HD void f12_reduce(F12& r, Fq* t, int top, const PairingConsts& k) {
  ZK_NOUNROLL for (int i = top; i >= 12; i--) {
    t[i - 6] = t[i - 6] + k.k18 * t[i];
    t[i - 12] = t[i - 12] - k.k82 * t[i];
  }
  ZK_NOUNROLL for (int i = 0; i < 12; i++) r.c[i] = t[i];
}
ZK_HD void f12_mul(F12& r, const F12& a, const F12& b, const PairingConsts& k) {   // r may alias a or b
  Fq t[23];
  ZK_NOUNROLL for (int i = 0; i < 23; i++) t[i] = Fq::zero();
  ZK_NOUNROLL for (int i = 0; i < 12; i++) {
    const Fq ai = a.c[i];
    ZK_NOUNROLL for (int j = 0; j < 12; j++) t[i + j] = t[i + j] + ai * b.c[j];
  }
  f12_reduce(r, t, 22, k);
}
ZK_HD void f12_sqr(F12& r, const F12& a, const PairingConsts& k) {
  Fq t[23];
  ZK_NOUNROLL for (int i = 0; i < 23; i++) t[i] = Fq::zero();
  ZK_NOUNROLL for (int i = 0; i < 12; i++) {
    const Fq ai = a.c[i];
    t[2 * i] = t[2 * i] + ai * ai;
    ZK_NOUNROLL for (int j = i + 1; j < 12; j++) { const Fq p = ai * a.c[j]; t[i + j] = t[i + j] + p.dbl(); }
  }
  f12_reduce(r, t, 22, k);
}
// f <- l * f for a line l = c0 + c1 w + c3 w^3 + c7 w^7 + c9 w^9
ZK_HD void f12_mul_line(F12& f, const Line& l, const PairingConsts& k) {
  Fq t[21];
  ZK_NOUNROLL for (int i = 0; i < 21; i++) t[i] = Fq::zero();
  ZK_NOUNROLL for (int s = 0; s < 5; s++) {
    const int d = s == 0 ? 0 : s == 1 ? 1 : s == 2 ? 3 : s == 3 ? 7 : 9;
    const Fq li = s == 0 ? l.c0 : s == 1 ? l.c1 : s == 2 ? l.c3 : s == 3 ? l.c7 : l.c9;
    ZK_NOUNROLL for (int j = 0; j < 12; j++) t[d + j] = t[d + j] + li * f.c[j];
  }
  f12_reduce(f, t, 20, k);
}
ZK_HD void f12_conj(F12& r, const F12& a) { ZK_NOUNROLL for (int i = 0; i < 12; i++) r.c[i] = (i & 1) ? a.c[i].neg() : a.c[i]; }
ZK_HD void f12_frob2(F12& r, const F12& a, const PairingConsts& k) { ZK_NOUNROLL for (int i = 0; i < 12; i++) r.c[i] = a.c[i] * k.zeta[i]; }
// r = a^h, h = (p^4 - p^2 + 1) / r_order (761 bits)
ZK_HD void f12_pow_h(F12& r, const F12& a, const PairingConsts& k) {
  constexpr uint32_t H[24] = {0xccdf42b1u, 0xe81bb482u, 0xf49c36d4u, 0x5abf5cc4u, 0x1da014fdu, 0xf1154e7eu, 0x87cdbacfu, 0xdcc7b44cu,
                              0x954bcf8au, 0xaaa441e3u, 0xd5095f23u, 0x6b887d56u, 0xf3fd90c6u, 0x79581e16u, 0xd189227du, 0x3b1b1355u,
                              0x61876f6bu, 0x4e529a58u, 0xd5b12278u, 0x6c0eb522u, 0x83177fafu, 0x331ec151u, 0x0b0759adu, 0x01baaa71u};
  r = a;   // bit 760 is the leading one
  ZK_NOUNROLL for (int i = 759; i >= 0; i--) {
    f12_sqr(r, r, k);
    if ((H[i >> 5] >> (i & 31)) & 1) f12_mul(r, r, a, k);
  }
}

// ------------------------------------------------------------------------------ curve checks and G2 line steps
ZK_HD bool g1_on_curve(const G1P& p, const PairingConsts& k) { if (p.inf) return true; return p.y.sqr() == p.x.sqr() * p.x + k.k3; }
ZK_HD bool g2_on_curve(const G2P& p, const PairingConsts& k) { if (p.inf) return true; return p.y.sqr() == p.x.sqr() * p.x + k.twist_b; }

// line through T and U (tangent when U == T) evaluated at P; T <- T + U.  false: vertical line (cannot occur for points of
// order r inside the loop; reported as an invalid proof)
ZK_HD bool line_step(G2P& T, const G2P& U, const G1P& P, Line& l, const PairingConsts& k) {
  Fq2 lam;
  if (T.x == U.x) {
    if (!(T.y == U.y) || T.y.is_zero()) return false;
    Fq2 n = T.x.sqr(); n = n.dbl() + n;
    lam = n * T.y.dbl().inv_gcd();
  } else {
    lam = (U.y - T.y) * (U.x - T.x).inv_gcd();
  }
  // l = -yP + (lam * xP) w + (yT - lam * xT) w^3, an Fq2 element (a + b u) embedded as (a - 9 b) + b w^6
  l.c0 = P.y.neg();
  const Fq a0 = lam.a * P.x, a1 = lam.b * P.x;
  l.c1 = a0 - k.k9 * a1; l.c7 = a1;
  const Fq2 m = T.y - lam * T.x;
  l.c3 = m.a - k.k9 * m.b; l.c9 = m.b;
  const Fq2 x3 = lam.sqr() - T.x - U.x;
  const Fq2 y3 = lam * (T.x - x3) - T.y;
  T.x = x3; T.y = y3;
  return true;
}
// f = Miller loop value of (Q, P) without the final exponentiation
ZK_HD bool miller(const G2P& Q, const G1P& P, F12& f, const PairingConsts& k) {
  f12_set_one(f);
  if (Q.inf || P.inf) return true;
  const uint64_t ate = 0x9d797039be763ba8ull;  // low 64 bits of 6x + 2 = 0x19d797039be763ba8 (bit 64 is the implicit leading one)
  G2P Q1; Q1.inf = 0; Q1.x = fq2_conj(Q.x) * k.g12; Q1.y = fq2_conj(Q.y) * k.g13;
  G2P Q2; Q2.inf = 0; Q2.x = fq2_conj(Q1.x) * k.g12; Q2.y = (fq2_conj(Q1.y) * k.g13).neg();
  G2P T = Q;
  Line l;
  // one loop over "steps": (doubling step of bit i, then the addition step of bit i if set), then the two Frobenius additions;
  // a single copy of the squaring, the line step and the sparse product in the code
  int i = 63, phase = 0;   // phase 0: doubling of bit i, 1: addition of bit i, 2: + Q1, 3: + Q2
  ZK_NOUNROLL for (;;) {
    if (phase == 0) f12_sqr(f, f, k);
    const G2P& U = phase == 0 ? T : phase == 1 ? Q : phase == 2 ? Q1 : Q2;
    G2P Uc = U;   // T is modified by the step
    if (!line_step(T, Uc, P, l, k)) return false;
    f12_mul_line(f, l, k);
    if (phase == 0) {
      if ((ate >> i) & 1) phase = 1; else if (i == 0) phase = 2; else i--;
    } else if (phase == 1) {
      if (i == 0) phase = 2; else { phase = 0; i--; }
    } else if (phase == 2) phase = 3;
    else break;
  }
  return true;
}

// ------------------------------------------------------------------------------ input decoding
ZK_HD bool canonical_lt(const uint32_t* v, bool fr) {
  ZK_NOUNROLL for (int i = 7; i >= 0; i--) {
    const uint32_t m = fr ? zk::FrP::mod(i) : zk::FqP::mod(i);
    if (v[i] < m) return true;
    if (v[i] > m) return false;
  }
  return false;
}
ZK_HD G1P g1_from_canonical(const uint32_t* w) {   // 16 words: x, y
  G1P r;
  ZK_UNROLL for (int i = 0; i < 8; i++) { r.x.v[i] = w[i]; r.y.v[i] = w[8 + i]; }
  r.inf = (r.x.is_zero() && r.y.is_zero()) ? 1u : 0u;
  r.x = r.x.to_mont(); r.y = r.y.to_mont();
  return r;
}
ZK_HD G2P g2_from_canonical(const uint32_t* w) {   // 32 words: x.c0, x.c1, y.c0, y.c1
  G2P r;
  ZK_UNROLL for (int i = 0; i < 8; i++) { r.x.a.v[i] = w[i]; r.x.b.v[i] = w[8 + i]; r.y.a.v[i] = w[16 + i]; r.y.b.v[i] = w[24 + i]; }
  r.inf = (r.x.is_zero() && r.y.is_zero()) ? 1u : 0u;
  r.x.a = r.x.a.to_mont(); r.x.b = r.x.b.to_mont(); r.y.a = r.y.a.to_mont(); r.y.b = r.y.b.to_mont();
  return r;
}
ZK_HD G1P g1_from_xyzz(const zk::G1Xyzz& p) {
  G1P r;
  zk::G1Affine a = zk::xyzz_to_affine(p);
  r.x = a.x; r.y = a.y; r.inf = p.is_inf() ? 1u : 0u;
  return r;
}
ZK_HD zk::G1Affine g1_to_affine(const G1P& p) { zk::G1Affine a; a.x = p.inf ? Fq::zero() : p.x; a.y = p.inf ? Fq::zero() : p.y; return a; }

// the proof-dependent part of the check, as the batch kernels split it:
//   F = miller(B, -A) * miller(gamma, vk_x) * miller(delta, C) * miller(beta, alpha);   valid <=> final_exp(F) == 1
ZK_HD void final_half(const F12& F, int side, F12& out, const PairingConsts& k) {   // side 0: conj(F), 1: F
  F12 x, y;
  ZK_NOUNROLL for (int i = 0; i < 12; i++) x.c[i] = (side == 0 && (i & 1)) ? F.c[i].neg() : F.c[i];   // conj(F) or F
  f12_frob2(y, x, k);
  f12_mul(y, y, x, k);
  f12_pow_h(out, y, k);
}
}  // namespace zkp
