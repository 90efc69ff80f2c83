// Groth16 verification arithmetic shared by the device batch verifier (kernels.cuh, section V1) and the single-proof host
// entry point (`snarkjs groth16 verify`, reference call sites tests/full_system_simulation.mjs:865-868,975-978,1116-1119;
// Server.verify*Proof, :848-1131).
//
// Optimal ate pairing on BN254 over Fq12 = Fq[w]/(w^12 - 18 w^6 + 82) (xi = w^6 = 9 + u), the flat basis of the oracle
// (the same representation the test oracle uses, so intermediate values can be compared), G2 arithmetic affine on the twist.  Every function is written once for host and device; the
// big ones are loops over small bodies (no unrolling, local arrays) so that a kernel holds ONE copy of the Fq12 product, the
// line step and the sparse product -- the device code has no calls with stack frames, only the Fq product leaf call.
//   * default Miller loop: homogeneous projective steps, accumulator in the tower view (miller_proj below); the affine loop
//     in the flat basis (miller) is the cross-check form (`ZKFL_VERIFY_FLAT=1`): sparse lines l = c0 + c1 w + c3 w^3 + c7 w^7 +
//     c9 w^9 -> 60 products instead of 144, squarings over the symmetric half: 78 products;
//   * final exponentiation: easy part + x-power chain in the tower view (final_exp_is_one below, the default); the first
//     version is kept for cross-checking (`ZKFL_VERIFY_FLAT=1`): f^((p^12-1)/r) == 1  <=>  (conj(f)/f)^((p^2+1) h) == 1,
//     h = (p^4 - p^2 + 1)/r,  <=>  (frob2(conj f) conj f)^h == (frob2(f) f)^h  -- no Fq12 inversion, two independent 761-bit
//     powers (two warps per 32 proofs on the device) instead of one 2816-bit power.  conj = Frobenius^6 flips the odd coefficients; frob2 = Frobenius^2
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
  Fq2 frob[3][6];     // frob[k-1][m] = xi^(m (p^k - 1)/6): Frobenius^k on the coefficient of w^m (tower view below)
  Fq half;            // 1/2
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
  // (p^k - 1) / 6 for k = 1, 2, 3
  static const uint32_t F1[8] = {0x2414d4e1u,0x34b01759u,0xe6bda1c2u,0xee9591c2u,0xc0403964u,0xf40d60f3u,0xd032f006u,0x0810b7bdu};
  static const uint32_t F2[16] = {0xb13a3c48u,0x348e0ec5u,0xd6fc7580u,0xc655abdcu,0xbcee7724u,0x0c62aec4u,0xe9adb5ccu,0x2b66c518u,
                                  0xb3767342u,0x5bd25464u,0x2e5e8e56u,0x72ac9638u,0xab36cdafu,0x0eef1294u,0x13b4ca9au,0x01864b74u};
  static const uint32_t F3[24] = {0xcbaeb4d9u,0x9ef31995u,0xec487080u,0xac3dad95u,0xf63bddf5u,0x8b33bea5u,0x984eeb22u,0x9fefedd1u,
                                  0x6ca1caa5u,0x6fea09beu,0x3f9c6113u,0x67d81a82u,0xd6398826u,0x00daed7bu,0xeb1f2783u,0x2667434cu,
                                  0x32525cfau,0xa0605a09u,0xd0fb6bfdu,0x3c036d4du,0x4083ea9du,0x88852038u,0xd72be447u,0x0049c712u};
  const Fq2 gam[3] = {fq2_pow_host(xi, F1, 8), fq2_pow_host(xi, F2, 16), fq2_pow_host(xi, F3, 24)};
  k.half = fq_small(2).inv();
  for (int q = 0; q < 3; q++) {
    k.frob[q][0] = Fq2::one();
    for (int m = 1; m < 6; m++) k.frob[q][m] = k.frob[q][m - 1] * gam[q];
  }
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

// ------------------------------------------------------------------------------ final exponentiation in the tower view
// The same field seen as Fq2[w]/(w^6 - xi): element = sum_m c[m] w^m, c[m] in Fq2.  From the flat basis (u = w^6 - 9):
// c[m] = (flat[m] + 9 flat[m+6]) + flat[m+6] u.  Products are degree-6 schoolbook over Fq2 (36 Fq2 = 108 Fq products, loops
// only), squarings the symmetric half (57), Frobenius^k = conj^k on the coefficients times frob[k-1][m], inversion through
// Fq6 = Fq2[v]/(v^3 - xi), v = w^2 (even / odd coefficients).  With those the classic split is affordable:
//   easy part  g = (conj(f) / f)^(p^2 + 1)   (one inversion),
//   hard part  g^(k h), k not divisible by r, by the x-power chain of Fuentes-Castaneda et al. (three 63-bit powers of the BN
//   parameter x = 0x44e992b44a6909f1, Frobenius^1,2,3): ~20 k Fq products instead of the 2 x 140 k of the two flat powers.
// The chain's exponent is checked symbolically in tests/test_oracle_pins.py; zkfl_debug_pairing_selftest checks the maps.
struct T12 { Fq2 c[6]; };
ZK_HD Fq2 mul_xi(const Fq2& x) {   // (9 + u) x
  const Fq a9 = x.a.dbl().dbl().dbl() + x.a, b9 = x.b.dbl().dbl().dbl() + x.b;
  Fq2 r; r.a = a9 - x.b; r.b = b9 + x.a; return r;
}
ZK_HD void t12_from_flat(T12& r, const F12& f, const PairingConsts& k) {
  ZK_NOUNROLL for (int m = 0; m < 6; m++) { r.c[m].b = f.c[m + 6]; r.c[m].a = f.c[m] + k.k9 * f.c[m + 6]; }
}
ZK_HD void t12_to_flat(F12& f, const T12& a, const PairingConsts& k) {
  ZK_NOUNROLL for (int m = 0; m < 6; m++) { f.c[m + 6] = a.c[m].b; f.c[m] = a.c[m].a - k.k9 * a.c[m].b; }
}
ZK_HD bool t12_is_one(const T12& a) {
  bool e = a.c[0] == Fq2::one();
  ZK_NOUNROLL for (int m = 1; m < 6; m++) e = e && a.c[m].is_zero();
  return e;
}
ZK_HD void t12_reduce(T12& r, Fq2* t) {   // w^6 = xi
  ZK_NOUNROLL for (int i = 10; i >= 6; i--) t[i - 6] = t[i - 6] + mul_xi(t[i]);
  ZK_NOUNROLL for (int i = 0; i < 6; i++) r.c[i] = t[i];
}
ZK_HD void t12_mul(T12& r, const T12& a, const T12& b) {   // r may alias a or b
  Fq2 t[11];
  ZK_NOUNROLL for (int i = 0; i < 11; i++) t[i] = Fq2::zero();
  ZK_NOUNROLL for (int i = 0; i < 6; i++) {
    const Fq2 ai = a.c[i];
    ZK_NOUNROLL for (int j = 0; j < 6; j++) t[i + j] = t[i + j] + ai * b.c[j];
  }
  t12_reduce(r, t);
}
ZK_HD void t12_sqr(T12& r, const T12& a) {
  Fq2 t[11];
  ZK_NOUNROLL for (int i = 0; i < 11; i++) t[i] = Fq2::zero();
  ZK_NOUNROLL for (int i = 0; i < 6; i++) {
    const Fq2 ai = a.c[i];
    t[2 * i] = t[2 * i] + ai.sqr();
    ZK_NOUNROLL for (int j = i + 1; j < 6; j++) t[i + j] = t[i + j] + (ai * a.c[j]).dbl();
  }
  t12_reduce(r, t);
}
ZK_HD void t12_conj(T12& r, const T12& a) { ZK_NOUNROLL for (int m = 0; m < 6; m++) r.c[m] = (m & 1) ? a.c[m].neg() : a.c[m]; }
ZK_HD void t12_frob(T12& r, const T12& a, int q, const PairingConsts& k) {   // Frobenius^q, q = 1, 2, 3
  ZK_NOUNROLL for (int m = 0; m < 6; m++) r.c[m] = ((q & 1) ? fq2_conj(a.c[m]) : a.c[m]) * k.frob[q - 1][m];
}
// a^-1 = (a0 - w a1) / (a0^2 - v a1^2), a = a0(v) + w a1(v) with a0, a1 in Fq6 (even / odd coefficients), v = w^2, v^3 = xi
ZK_HD void f6_mul(Fq2* r, const Fq2* x, const Fq2* y) {   // r must not alias
  Fq2 t[5];
  ZK_NOUNROLL for (int i = 0; i < 5; i++) t[i] = Fq2::zero();
  ZK_NOUNROLL for (int i = 0; i < 3; i++) ZK_NOUNROLL for (int j = 0; j < 3; j++) t[i + j] = t[i + j] + x[i] * y[j];
  r[0] = t[0] + mul_xi(t[3]); r[1] = t[1] + mul_xi(t[4]); r[2] = t[2];
}
ZK_HD void t12_inv(T12& r, const T12& a) {
  Fq2 a0[3], a1[3], s0[3], s1[3], d[3];
  ZK_NOUNROLL for (int j = 0; j < 3; j++) { a0[j] = a.c[2 * j]; a1[j] = a.c[2 * j + 1]; }
  f6_mul(s0, a0, a0);
  f6_mul(s1, a1, a1);
  // d = a0^2 - v a1^2;  v (s1_0, s1_1, s1_2) = (xi s1_2, s1_0, s1_1)
  d[0] = s0[0] - mul_xi(s1[2]); d[1] = s0[1] - s1[0]; d[2] = s0[2] - s1[1];
  // d^-1 in Fq6
  const Fq2 t0 = d[0].sqr() - mul_xi(d[1] * d[2]);
  const Fq2 t1 = mul_xi(d[2].sqr()) - d[0] * d[1];
  const Fq2 t2 = d[1].sqr() - d[0] * d[2];
  const Fq2 n = (d[0] * t0 + mul_xi(d[2] * t1 + d[1] * t2)).inv_gcd();
  Fq2 di[3]; di[0] = t0 * n; di[1] = t1 * n; di[2] = t2 * n;
  f6_mul(s0, a0, di);
  f6_mul(s1, a1, di);
  ZK_NOUNROLL for (int j = 0; j < 3; j++) { r.c[2 * j] = s0[j]; r.c[2 * j + 1] = s1[j].neg(); }
}
ZK_HD void t12_exp_neg_x(T12& r, const T12& a) {   // conj(a^x): a^(-x) inside the cyclotomic subgroup
  const uint64_t x = 0x44e992b44a6909f1ull;      // bit 62 is the leading one
  T12 acc = a;
  ZK_NOUNROLL for (int i = 61; i >= 0; i--) {
    t12_sqr(acc, acc);
    if ((x >> i) & 1) t12_mul(acc, acc, a);
  }
  t12_conj(r, acc);
}
// out = f^((p^12 - 1)/r * k)  (k not divisible by r)
ZK_HD void final_exp_value(const F12& flat, T12& y3, const PairingConsts& k) {
  T12 f, g, y0, y1, y2, y4, y6, t;
  t12_from_flat(f, flat, k);
  // easy part
  t12_inv(t, f);
  t12_conj(g, f);
  t12_mul(g, g, t);            // f^(p^6 - 1)
  t12_frob(t, g, 2, k);
  t12_mul(g, t, g);            // ^(p^2 + 1): g is in the cyclotomic subgroup, where conj = inverse
  // hard part
  t12_exp_neg_x(y0, g);
  t12_sqr(y1, y0);
  t12_sqr(y2, y1);
  t12_mul(y3, y2, y1);
  t12_exp_neg_x(y4, y3);
  t12_sqr(t, y4);              // y5
  t12_exp_neg_x(y6, t);
  t12_conj(y3, y3);
  t12_conj(y6, y6);
  t12_mul(y6, y6, y4);         // y7
  t12_mul(y6, y6, y3);         // y8
  t12_mul(y2, y6, y1);         // y9
  t12_mul(y3, y6, y4);         // y10
  t12_mul(y3, y3, g);          // y11
  t12_frob(t, y2, 1, k);       // y12
  t12_mul(y3, t, y3);          // y13
  t12_frob(t, y6, 2, k);
  t12_mul(y3, t, y3);          // y14
  t12_conj(t, g);
  t12_mul(t, t, y2);           // r^-1 y9
  t12_frob(y0, t, 3, k);       // y15
  t12_mul(y3, y0, y3);         // y16
}
ZK_HD bool final_exp_is_one(const F12& flat, const PairingConsts& k) {
  T12 v;
  final_exp_value(flat, v, k);
  return t12_is_one(v);
}

// ------------------------------------------------------------------------------ Miller loop without inversions
// Homogeneous projective steps on the twist (Costello-Lange-Naehrig formulas as commonly implemented for D-type twists) and
// the accumulator in the tower view: a doubling step is 57 (square) + 28 (point and line) + 58 (sparse product) Fq products
// against ~290 with the affine line step, whose slope inversion alone is a third of the loop.  The lines differ from the
// affine ones by Fq2 factors, which the final exponentiation removes: the Miller values differ, the pairing values do not
// (zkfl_debug_pairing_selftest compares the exponentiated values of both loops).
struct G2H { Fq2 x, y, z; };
struct LineT { Fq2 c0, c1, c3; };   // c0 + c1 w + c3 w^3, before the evaluation at P (c0 *= yP, c1 *= xP)
ZK_HD Fq2 fq2_scale(const Fq2& x, const Fq& s) { Fq2 r; r.a = x.a * s; r.b = x.b * s; return r; }
ZK_HD void proj_dbl_step(G2H& r, LineT& l, const PairingConsts& k) {
  const Fq2 a = fq2_scale(r.x * r.y, k.half);
  const Fq2 b = r.y.sqr(), c = r.z.sqr();
  const Fq2 e = k.twist_b * (c.dbl() + c);
  const Fq2 f = e.dbl() + e;
  const Fq2 g = fq2_scale(b + f, k.half);
  const Fq2 h = (r.y + r.z).sqr() - (b + c);
  const Fq2 j = r.x.sqr();
  const Fq2 e2 = e.sqr();
  l.c0 = h.neg(); l.c1 = j.dbl() + j; l.c3 = e - b;
  r.x = a * (b - f);
  r.y = g.sqr() - (e2.dbl() + e2);
  r.z = b * h;
}
ZK_HD void proj_add_step(G2H& r, const G2P& q, LineT& l) {
  const Fq2 theta = r.y - q.y * r.z, lambda = r.x - q.x * r.z;
  const Fq2 c = theta.sqr(), d = lambda.sqr();
  const Fq2 e = lambda * d, f = r.z * c, g = r.x * d;
  const Fq2 h = e + f - g.dbl();
  l.c0 = lambda; l.c1 = theta.neg(); l.c3 = theta * q.x - lambda * q.y;
  r.x = lambda * h;
  r.y = theta * (g - h) - e * r.y;
  r.z = r.z * e;
}
// f <- f * (c0 yP + c1 xP w + c3 w^3)
ZK_HD void t12_mul_line(T12& f, const LineT& l, const G1P& P) {
  const Fq2 c0 = fq2_scale(l.c0, P.y), c1 = fq2_scale(l.c1, P.x);
  Fq2 t[11];
  ZK_NOUNROLL for (int i = 0; i < 11; i++) t[i] = Fq2::zero();
  ZK_NOUNROLL for (int s = 0; s < 3; s++) {
    const int d = s == 0 ? 0 : s == 1 ? 1 : 3;
    const Fq2 li = s == 0 ? c0 : s == 1 ? c1 : l.c3;
    ZK_NOUNROLL for (int j = 0; j < 6; j++) t[d + j] = t[d + j] + li * f.c[j];
  }
  t12_reduce(f, t);
}
ZK_HD void miller_proj(const G2P& Q, const G1P& P, F12& out, const PairingConsts& k) {
  T12 f;
  ZK_NOUNROLL for (int m = 0; m < 6; m++) f.c[m] = Fq2::zero();
  f.c[0] = Fq2::one();
  if (!(Q.inf || P.inf)) {
    const uint64_t ate = 0x9d797039be763ba8ull;  // low 64 bits of 6x + 2, bit 64 is the implicit leading one
    G2P Q1; Q1.inf = 0; Q1.x = fq2_conj(Q.x) * k.g12; Q1.y = fq2_conj(Q.y) * k.g13;
    G2P Q2; Q2.inf = 0; Q2.x = fq2_conj(Q1.x) * k.g12; Q2.y = (fq2_conj(Q1.y) * k.g13).neg();
    G2H R; R.x = Q.x; R.y = Q.y; R.z = Fq2::one();
    LineT l;
    int i = 63, phase = 0;   // phase 0: doubling of bit i, 1: addition of bit i, 2: + Q1, 3: + Q2
    ZK_NOUNROLL for (;;) {
      if (phase == 0) { t12_sqr(f, f); proj_dbl_step(R, l, k); }
      else proj_add_step(R, phase == 1 ? Q : phase == 2 ? Q1 : Q2, l);
      t12_mul_line(f, l, P);
      if (phase == 0) {
        if ((ate >> i) & 1) phase = 1; else if (i == 0) phase = 2; else i--;
      } else if (phase == 1) {
        if (i == 0) phase = 2; else { phase = 0; i--; }
      } else if (phase == 2) phase = 3;
      else break;
    }
  }
  t12_to_flat(out, f, k);
}

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
