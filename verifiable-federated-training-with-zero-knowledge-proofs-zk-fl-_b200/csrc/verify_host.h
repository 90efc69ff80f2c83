// Host-side Groth16 verifier for the drop-in boundary (`snarkjs groth16 verify`, reference call sites
// tests/full_system_simulation.mjs:865-868,975-978,1116-1119).  SURVEY section 2.4 row V1 keeps verification on the
// CPU (three pairings per proof, O(1) work); it uses the library's own BN254 field code on its host path.
// Optimal ate pairing over Fq12 = Fq[w]/(w^12 - 18 w^6 + 82) (u = w^6 - 9), G2 arithmetic in affine Fq2 on the twist,
// final exponentiation by plain square-and-multiply with the fixed exponent (q^12 - 1)/r.
#pragma once
#include "bn254.cuh"

namespace zkv {
using zk::Fq; using zk::Fq2; using zk::Fr;

struct F12 { Fq c[12]; };
static inline Fq fq_small(uint32_t v) { Fq r = Fq::zero(); r.v[0] = v; return r.to_mont(); }
static inline F12 f12_one() { F12 r; for (auto& x : r.c) x = Fq::zero(); r.c[0] = Fq::one(); return r; }
static inline bool f12_is_one(const F12& a) { if (!(a.c[0] == Fq::one())) return false; for (int i = 1; i < 12; i++) if (!a.c[i].is_zero()) return false; return true; }
static F12 f12_mul(const F12& a, const F12& b) {
  static const Fq k18 = fq_small(18), k82 = fq_small(82);
  Fq t[23];
  for (auto& x : t) x = Fq::zero();
  for (int i = 0; i < 12; i++) {
    if (a.c[i].is_zero()) continue;
    for (int j = 0; j < 12; j++) t[i + j] = t[i + j] + a.c[i] * b.c[j];
  }
  for (int i = 22; i >= 12; i--) {  // w^12 = 18 w^6 - 82
    if (t[i].is_zero()) continue;
    t[i - 6] = t[i - 6] + k18 * t[i];
    t[i - 12] = t[i - 12] - k82 * t[i];
  }
  F12 r; for (int i = 0; i < 12; i++) r.c[i] = t[i];
  return r;
}
// embeds (a + b u) * w^k, k < 6
static void f12_add_fq2(F12& r, const Fq2& x, int k) {
  static const Fq k9 = fq_small(9);
  r.c[k] = r.c[k] + (x.a - k9 * x.b);
  r.c[k + 6] = r.c[k + 6] + x.b;
}
static Fq2 fq2_pow(Fq2 base, const uint32_t* e, int nwords) {
  Fq2 r = Fq2::one();
  for (int i = nwords * 32 - 1; i >= 0; i--) { r = r.sqr(); if ((e[i >> 5] >> (i & 31)) & 1) r = r * base; }
  return r;
}
static Fq2 fq2_conj(const Fq2& x) { Fq2 r; r.a = x.a; r.b = x.b.neg(); return r; }

struct G1 { Fq x, y; bool inf; };
struct G2 { Fq2 x, y; bool inf; };
static bool g1_on_curve(const G1& p) { if (p.inf) return true; return p.y.sqr() == p.x.sqr() * p.x + fq_small(3); }
static Fq2 twist_b() { Fq2 xi; xi.a = fq_small(9); xi.b = fq_small(1); Fq2 three; three.a = fq_small(3); three.b = Fq::zero(); return three * xi.inv(); }
static bool g2_on_curve(const G2& p) { if (p.inf) return true; return p.y.sqr() == p.x.sqr() * p.x + twist_b(); }

// line through T (and U, or tangent when U == T) evaluated at P; returns l and advances T <- T + U
static bool line_step(G2& T, const G2& U, const G1& P, F12& l) {
  Fq2 lam;
  if (T.x == U.x) {
    if (!(T.y == U.y) || T.y.is_zero()) return false;  // vertical line: cannot occur for points of order r in the loop
    Fq2 n = T.x.sqr(); n = n.dbl() + n;
    lam = n * T.y.dbl().inv();
  } else {
    lam = (U.y - T.y) * (U.x - T.x).inv();
  }
  // l = -yP + (lam * xP) w + (yT - lam * xT) w^3
  for (auto& x : l.c) x = Fq::zero();
  l.c[0] = P.y.neg();
  Fq2 a; a.a = lam.a * P.x; a.b = lam.b * P.x;
  f12_add_fq2(l, a, 1);
  f12_add_fq2(l, T.y - lam * T.x, 3);
  Fq2 x3 = lam.sqr() - T.x - U.x;
  Fq2 y3 = lam * (T.x - x3) - T.y;
  T.x = x3; T.y = y3;
  return true;
}
static bool miller(const G2& Q, const G1& P, F12& f) {
  f = f12_one();
  if (Q.inf || P.inf) return true;
  const uint64_t ate = 0x9d797039be763ba8ull;  // low 64 bits of 6x + 2 = 0x19d797039be763ba8 (bit 64 is the implicit leading one)
  G2 T = Q; F12 l;
  for (int i = 63; i >= 0; i--) {
    f = f12_mul(f, f);
    if (!line_step(T, T, P, l)) return false;
    f = f12_mul(f, l);
    if ((ate >> i) & 1) { if (!line_step(T, Q, P, l)) return false; f = f12_mul(f, l); }
  }
  static const uint32_t E3[8] = {0x4829a9c2u, 0x69602eb2u, 0xcd7b4384u, 0xdd2b2385u, 0x808072c9u, 0xe81ac1e7u, 0xa065e00du, 0x10216f7bu};
  static const uint32_t E2[8] = {0x6c3e7ea3u, 0x9e10460bu, 0xb438e546u, 0xcbc0b548u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
  Fq2 xi; xi.a = fq_small(9); xi.b = fq_small(1);
  const Fq2 g12 = fq2_pow(xi, E3, 8), g13 = fq2_pow(xi, E2, 8);
  G2 Q1; Q1.inf = false; Q1.x = fq2_conj(Q.x) * g12; Q1.y = fq2_conj(Q.y) * g13;
  G2 Q2; Q2.inf = false; Q2.x = fq2_conj(Q1.x) * g12; Q2.y = (fq2_conj(Q1.y) * g13).neg();
  if (!line_step(T, Q1, P, l)) return false;
  f = f12_mul(f, l);
  if (!line_step(T, Q2, P, l)) return false;
  f = f12_mul(f, l);
  return true;
}
static F12 final_exp(const F12& f) {
  static const uint32_t E[88] = {0xca86f120u,0x86964b64u,0xe54523a4u,0x40a4efb7u,0x96e84abbu,0x837fa978u,0xb9b2b918u,0x361102b6u,0xf35692dau,0xc0de81deu,0xa6c3c760u,0xbe04c7e8u,0xd570bb7fu,0xd766f9c9u,0x83561841u,0xc230974du,0xc3be69a3u,0x5bba1668u,0x10526294u,0x7f3811c4u,0xdadda71cu,0x29baee7du,0x145da900u,0xbf813b8du,0x423f9a2cu,0x641bbadfu,0x44eacc5eu,0xa80bb4eau,0x14fde37cu,0xcd656648u,0x580291d2u,0x4a0364b9u,0x0826f0ddu,0xee93dfb1u,0xc5514724u,0x6b42db8du,0x0b0f3785u,0xbb10cf43u,0x6f804216u,0x40494e40u,0xacf3aafbu,0x55cfe107u,0xe0ebae87u,0x2088ec80u,0x11a337a0u,0x846a3ed0u,0x1e3a5195u,0x48a45a4au,0xdfc50e16u,0xe5664568u,0x4c0cc4ebu,0xab6a4129u,0xd268c7dau,0x82d0d602u,0xed3cc48au,0x6668449au,0xb2015dfcu,0x5062cd0fu,0xb1ddb3d1u,0x7f2940a8u,0x2a226448u,0x77f5b63au,0x61e443aeu,0xfef07813u,0x88d5c6c8u,0xf977870eu,0x1f676baau,0x790364a6u,0xceaddea3u,0x5887e72eu,0xa09a1b70u,0x1377e563u,0x1bd8c3b2u,0x0c54efeeu,0xd524d8f7u,0x3ec3d15au,0xb2383a5du,0xdaf15466u,0xbb94fec0u,0xe1e30a73u,0x5f3f7be2u,0x6a1c7101u,0x6369b1ffu,0x842d43bfu,0x107d20bcu,0x20fddadfu,0x4b6dc970u,0x0000002fu};
  F12 r = f12_one();
  bool started = false;
  for (int i = 88 * 32 - 1; i >= 0; i--) {
    if (started) r = f12_mul(r, r);
    if ((E[i >> 5] >> (i & 31)) & 1) { r = started ? f12_mul(r, f) : f; started = true; }
  }
  return r;
}

static G1 g1_from_canonical(const uint8_t* p) {
  G1 r; memcpy(r.x.v, p, 32); memcpy(r.y.v, p + 32, 32);
  r.inf = r.x.is_zero() && r.y.is_zero();
  r.x = r.x.to_mont(); r.y = r.y.to_mont();
  return r;
}
static G2 g2_from_canonical(const uint8_t* p) {
  G2 r; memcpy(r.x.a.v, p, 32); memcpy(r.x.b.v, p + 32, 32); memcpy(r.y.a.v, p + 64, 32); memcpy(r.y.b.v, p + 96, 32);
  r.inf = r.x.is_zero() && r.y.is_zero();
  r.x.a = r.x.a.to_mont(); r.x.b = r.x.b.to_mont(); r.y.a = r.y.a.to_mont(); r.y.b = r.y.b.to_mont();
  return r;
}
static G1 g1_add(const G1& a, const G1& b) {
  if (a.inf) return b;
  if (b.inf) return a;
  Fq lam;
  if (a.x == b.x) {
    if (!(a.y == b.y) || a.y.is_zero()) { G1 r; r.inf = true; r.x = Fq::zero(); r.y = Fq::zero(); return r; }
    Fq n = a.x.sqr(); n = n.dbl() + n; lam = n * a.y.dbl().inv();
  } else lam = (b.y - a.y) * (b.x - a.x).inv();
  G1 r; r.inf = false; r.x = lam.sqr() - a.x - b.x; r.y = lam * (a.x - r.x) - a.y;
  return r;
}
static G1 g1_mul(const G1& p, const uint32_t* k) {
  G1 r; r.inf = true; r.x = Fq::zero(); r.y = Fq::zero();
  for (int i = 255; i >= 0; i--) { r = g1_add(r, r); if ((k[i >> 5] >> (i & 31)) & 1) r = g1_add(r, p); }
  return r;
}
static bool canonical_lt(const uint8_t* p, uint32_t (*mod)(int)) {
  uint32_t v[8]; memcpy(v, p, 32);
  for (int i = 7; i >= 0; i--) { if (v[i] < mod(i)) return true; if (v[i] > mod(i)) return false; }
  return false;
}
static uint32_t fq_mod(int i) { return zk::FqP::mod(i); }
static uint32_t fr_mod(int i) { return zk::FrP::mod(i); }

// returns 1 = valid, 0 = invalid, negative = malformed input.  All points affine canonical little-endian.
static int groth16_verify(const uint8_t alpha1[64], const uint8_t beta2[128], const uint8_t gamma2[128], const uint8_t delta2[128],
                          const uint8_t* ic /* (l+1) x 64 */, const uint8_t* publics /* l x 32 */, uint32_t l, const uint8_t proof[256]) {
  for (int i = 0; i < 8; i++) if (!canonical_lt(proof + 32 * i, fq_mod)) return 0;
  for (uint32_t i = 0; i < l; i++) if (!canonical_lt(publics + 32 * i, fr_mod)) return 0;  // snarkjs: public signals must be < r
  G1 A = g1_from_canonical(proof), C = g1_from_canonical(proof + 192);
  G2 Bp = g2_from_canonical(proof + 64);
  if (!g1_on_curve(A) || !g1_on_curve(C) || !g2_on_curve(Bp)) return 0;
  G1 vkx = g1_from_canonical(ic);
  for (uint32_t i = 0; i < l; i++) {
    uint32_t k[8]; memcpy(k, publics + 32 * i, 32);
    vkx = g1_add(vkx, g1_mul(g1_from_canonical(ic + 64 * (size_t)(i + 1)), k));
  }
  G1 nA = A; nA.y = A.y.neg();
  F12 f = f12_one(), t;
  const G1 g1s[4] = {nA, g1_from_canonical(alpha1), vkx, C};
  const G2 g2s[4] = {Bp, g2_from_canonical(beta2), g2_from_canonical(gamma2), g2_from_canonical(delta2)};
  for (int i = 0; i < 4; i++) {
    if (!miller(g2s[i], g1s[i], t)) return 0;
    f = f12_mul(f, t);
  }
  return f12_is_one(final_exp(f)) ? 1 : 0;
}
}  // namespace zkv
