// Lane-cooperative Fq12 arithmetic for the batch Groth16 verifier (Server.verify*Proof,
// tests/full_system_simulation.mjs:848-1131).  The single-thread forms of pairing.cuh are serial chains of 10-20 k Montgomery
// products per proof: with a few thousand proofs the GPU is latency-bound (3 072 proofs: 24 ms, flat in the batch size).  Here an
// element of Fq12 = Fq2[w]/(w^6 - xi) is spread over a GROUP OF 8 LANES: lane k < 6 holds the coefficient of w^k (lanes 6, 7 mirror
// lanes 0, 1 so that every shuffle stays inside an aligned group of 8), a product is 6 Fq2 products per lane (36 in one thread), a
// squaring 4 (the symmetric half, pairs taken from a per-lane table), a sparse line product 3.  Operands travel by warp shuffles
// (32 words per Fq2 pair against ~900 instructions per Fq2 product).
// The host emulation has no lanes: there ONE emulated thread holds all six coefficients (CNL = 6) and a "shuffle" is an array
// access -- the same source runs both ways, which is what lets the CPU test-suite cover this file.
#pragma once
#include "pairing.cuh"

namespace zkp {

#ifdef ZKFL_EMUL
constexpr int CNL = 6;
#else
constexpr int CNL = 1;
#endif
struct C12 { Fq2 c[CNL]; };

// lane context: k = coefficient index of this lane (device) / unused (emulation)
struct CoopLane {
  int k;
  uint32_t mask;   // the 8 lanes of this group: shuffles synchronise the group only, groups of one warp may diverge from each other
#ifndef ZKFL_EMUL
  __device__ __forceinline__ CoopLane() { const int l = (int)(threadIdx.x & 7u); k = l < 6 ? l : l - 6; mask = 0xFFu << (threadIdx.x & 24u); }
#else
  CoopLane() : k(0), mask(0) {}
#endif
};
// iterate over the coefficients this thread owns: l = slot in C12::c, kk = coefficient index
#define ZK_CLANES(L, l, kk) for (int l = 0, kk = (CNL == 1 ? (L).k : 0); l < CNL; l++, kk++)

#ifndef ZKFL_EMUL
__device__ __forceinline__ Fq2 coop_get(const C12& a, int src, const CoopLane& L) {   // coefficient `src` of the group's element
  Fq2 r;
  ZK_UNROLL for (int i = 0; i < 8; i++) {
    r.a.v[i] = __shfl_sync(L.mask, a.c[0].a.v[i], src, 8);
    r.b.v[i] = __shfl_sync(L.mask, a.c[0].b.v[i], src, 8);
  }
  return r;
}
__device__ __forceinline__ bool coop_all(bool v, const CoopLane& L) {   // AND over the group of 8
  uint32_t x = v ? 1u : 0u;
  x &= __shfl_xor_sync(L.mask, x, 1, 8);
  x &= __shfl_xor_sync(L.mask, x, 2, 8);
  x &= __shfl_xor_sync(L.mask, x, 4, 8);
  return x != 0;
}
#else
inline Fq2 coop_get(const C12& a, int src, const CoopLane&) { return a.c[src]; }
#endif

ZK_D void c12_set_one(C12& r, const CoopLane& L) { ZK_CLANES(L, l, k) r.c[l] = k == 0 ? Fq2::one() : Fq2::zero(); }
// global memory <-> lanes: T12 layout (six Fq2), lane k moves coefficient k
ZK_D void c12_load(C12& r, const T12* p, const CoopLane& L) { ZK_CLANES(L, l, k) r.c[l] = p->c[k]; }
ZK_D void c12_store(T12* p, const C12& a, const CoopLane& L, bool active) { ZK_CLANES(L, l, k) if (active) p->c[k] = a.c[l]; }

// r = a * b   (r may alias a or b)
ZK_D void c12_mul(C12& r, const C12& a, const C12& b, const CoopLane& L) {
  C12 lo, hi;
  ZK_CLANES(L, l, k) { lo.c[l] = Fq2::zero(); hi.c[l] = Fq2::zero(); }
  ZK_NOUNROLL for (int s = 0; s < 6; s++) {
    ZK_CLANES(L, l, k) {
      const Fq2 as = coop_get(a, s, L);
      const int src = k - s < 0 ? k - s + 6 : k - s;
      const Fq2 bs = coop_get(b, src, L);
      const Fq2 p = as * bs;
      const bool wrap = s > k;                       // s + src = k + 6: w^6 = xi
      lo.c[l] = wrap ? lo.c[l] : lo.c[l] + p;
      hi.c[l] = wrap ? hi.c[l] + p : hi.c[l];
    }
  }
  ZK_CLANES(L, l, k) r.c[l] = lo.c[l] + mul_xi(hi.c[l]);
}
// r = a^2: coefficient k = sum over unordered pairs {i, j}, i + j = k (mod 6); four slots per lane, packed per-lane tables
// (nibble k of word `slot`): I, J = the pair, XI = i + j >= 6, DB = i != j (counted twice), VA = slot used (odd k: three pairs)
ZK_D void c12_sqr(C12& r, const C12& a, const CoopLane& L) {
  // pairs per (k, slot):   k = 0: (0,0) (3,3)x (1,5)x2 (2,4)x2     k = 1: (0,1)2 (2,5)x2 (3,4)x2 -
  //  (x: times xi,        k = 2: (1,1) (4,4)x (0,2)2 (3,5)x2      k = 3: (0,3)2 (1,2)2 (4,5)x2 -
  //   2: doubled)         k = 4: (2,2) (5,5)x (0,4)2 (1,3)2       k = 5: (0,5)2 (1,4)2 (2,3)2 -
  C12 lo, hi;
  ZK_CLANES(L, l, k) { lo.c[l] = Fq2::zero(); hi.c[l] = Fq2::zero(); }
  ZK_UNROLL for (int s = 0; s < 4; s++) {
    const uint32_t pi = s == 0 ? 0x020100u : s == 1 ? 0x151423u : s == 2 ? 0x204031u : 0x010302u;
    const uint32_t pj = s == 0 ? 0x523110u : s == 1 ? 0x452453u : s == 2 ? 0x345245u : 0x030504u;
    const uint32_t px = s == 0 ? 0x00u : s == 1 ? 0x17u : s == 2 ? 0x0Bu : 0x05u;
    const uint32_t pd = s == 0 ? 0x2Au : s == 1 ? 0x2Au : s == 2 ? 0x3Fu : 0x15u;
    const uint32_t pv = s == 3 ? 0x15u : 0x3Fu;
    ZK_CLANES(L, l, k) {
      const Fq2 x = coop_get(a, (int)((pi >> (4 * k)) & 15u), L), y = coop_get(a, (int)((pj >> (4 * k)) & 15u), L);
      Fq2 p = x * y;
      if ((pd >> k) & 1u) p = p.dbl();
      if (!((pv >> k) & 1u)) p = Fq2::zero();
      const bool wrap = ((px >> k) & 1u) != 0;
      lo.c[l] = wrap ? lo.c[l] : lo.c[l] + p;
      hi.c[l] = wrap ? hi.c[l] + p : hi.c[l];
    }
  }
  ZK_CLANES(L, l, k) r.c[l] = lo.c[l] + mul_xi(hi.c[l]);
}
// f <- f * (c0 + c1 w + c3 w^3): coefficient k = c0 f_k + c1 f_(k-1) + c3 f_(k-3), indices mod 6, xi on wrap-around
ZK_D void c12_mul_line(C12& f, const Fq2& c0, const Fq2& c1, const Fq2& c3, const CoopLane& L) {
  C12 r;
  ZK_CLANES(L, l, k) {
    const Fq2 f1 = coop_get(f, k >= 1 ? k - 1 : k + 5, L), f3 = coop_get(f, k >= 3 ? k - 3 : k + 3, L);
    const Fq2 lo = c0 * coop_get(f, k, L);
    const Fq2 t1 = c1 * f1, t3 = c3 * f3;
    // wrapped terms take xi: gather them first, one mul_xi
    Fq2 hi = Fq2::zero(), acc = lo;
    if (k < 1) hi = hi + t1; else acc = acc + t1;
    if (k < 3) hi = hi + t3; else acc = acc + t3;
    r.c[l] = acc + mul_xi(hi);
  }
  ZK_CLANES(L, l, k) f.c[l] = r.c[l];
}
ZK_D void c12_conj(C12& r, const C12& a, const CoopLane& L) { ZK_CLANES(L, l, k) r.c[l] = (k & 1) ? a.c[l].neg() : a.c[l]; }
ZK_D void c12_frob(C12& r, const C12& a, int q, const PairingConsts& K, const CoopLane& L) {   // Frobenius^q, q = 1, 2, 3
  ZK_CLANES(L, l, k) r.c[l] = ((q & 1) ? fq2_conj(a.c[l]) : a.c[l]) * K.frob[q - 1][k];
}
// inversion: every lane gathers the whole element and runs the (short) tower inversion itself -- once per final exponentiation
ZK_D void c12_inv(C12& r, const C12& a, const CoopLane& L) {
  T12 t, u;
  ZK_NOUNROLL for (int m = 0; m < 6; m++) t.c[m] = coop_get(a, m, L);
  t12_inv(u, t);
  ZK_CLANES(L, l, k) r.c[l] = u.c[k];
}
ZK_D bool c12_is_one(const C12& a, const CoopLane& L) {
  bool e = true;
  ZK_CLANES(L, l, k) e = e && (k == 0 ? a.c[l] == Fq2::one() : a.c[l].is_zero());
#ifndef ZKFL_EMUL
  e = coop_all(e, L);
#endif
  return e;
}
ZK_D void c12_exp_neg_x(C12& r, const C12& a, const CoopLane& L) {   // conj(a^x): a^(-x) inside the cyclotomic subgroup
  const uint64_t x = 0x44e992b44a6909f1ull;      // bit 62 is the leading one
  C12 acc = a;
  ZK_NOUNROLL for (int i = 61; i >= 0; i--) {
    c12_sqr(acc, acc, L);
    if ((x >> i) & 1) c12_mul(acc, acc, a, L);
  }
  c12_conj(r, acc, L);
}
// the chain of final_exp_value (pairing.cuh), on lanes: true <=> f^((p^12 - 1)/r) == 1
ZK_D bool c12_final_exp_is_one(const C12& f, const PairingConsts& K, const CoopLane& L) {
  C12 g, t, y0, y1, y2, y3, y4, y6;
  c12_inv(t, f, L);
  c12_conj(g, f, L);
  c12_mul(g, g, t, L);             // f^(p^6 - 1)
  c12_frob(t, g, 2, K, L);
  c12_mul(g, t, g, L);             // ^(p^2 + 1)
  c12_exp_neg_x(y0, g, L);
  c12_sqr(y1, y0, L);
  c12_sqr(y2, y1, L);
  c12_mul(y3, y2, y1, L);
  c12_exp_neg_x(y4, y3, L);
  c12_sqr(t, y4, L);               // y5
  c12_exp_neg_x(y6, t, L);
  c12_conj(y3, y3, L);
  c12_conj(y6, y6, L);
  c12_mul(y6, y6, y4, L);          // y7
  c12_mul(y6, y6, y3, L);          // y8
  c12_mul(y2, y6, y1, L);          // y9
  c12_mul(y3, y6, y4, L);          // y10
  c12_mul(y3, y3, g, L);           // y11
  c12_frob(t, y2, 1, K, L);        // y12
  c12_mul(y3, t, y3, L);           // y13
  c12_frob(t, y6, 2, K, L);
  c12_mul(y3, t, y3, L);           // y14
  c12_conj(t, g, L);
  c12_mul(t, t, y2, L);            // r^-1 y9
  c12_frob(y0, t, 3, K, L);        // y15
  c12_mul(y3, y0, y3, L);          // y16
  return c12_is_one(y3, L);
}

// ------------------------------------------------------------------------------ Miller loop split in two
// (1) the G2 side alone (one thread per G2 point): the projective steps of miller_proj, every line stored UNSCALED
//     (c0, c1, c3 before the evaluation at P) -- independent of the G1 point, so the tables of the key's fixed points
//     (beta, gamma, delta) serve every proof; (2) the accumulator f <- f^2 * line on a lane group, reading the table.
constexpr int kMillerSteps = 64 + 36 + 2;   // doublings + additions of the set bits of 6x + 2 (below its leading one) + Q1, Q2
struct LineRec { Fq2 c0, c1, c3; };
// schedule shared by both halves: step -> is it a doubling step (f is squared first)?
struct MillerSched {
  int i, phase;
  ZK_HD MillerSched() : i(63), phase(0) {}
  ZK_HD bool is_dbl() const { return phase == 0; }
  ZK_HD bool advance() {   // false after the last step
    const uint64_t ate = 0x9d797039be763ba8ull;
    if (phase == 0) {
      if ((ate >> i) & 1) phase = 1; else if (i == 0) phase = 2; else i--;
    } else if (phase == 1) {
      if (i == 0) phase = 2; else { phase = 0; i--; }
    } else if (phase == 2) phase = 3;
    else return false;
    return true;
  }
};
ZK_HD void miller_lines(const G2P& Q, LineRec* __restrict__ out, const PairingConsts& k) {
  G2P Q1; Q1.inf = 0; Q1.x = fq2_conj(Q.x) * k.g12; Q1.y = fq2_conj(Q.y) * k.g13;
  G2P Q2; Q2.inf = 0; Q2.x = fq2_conj(Q1.x) * k.g12; Q2.y = (fq2_conj(Q1.y) * k.g13).neg();
  G2H R; R.x = Q.x; R.y = Q.y; R.z = Fq2::one();
  LineT l;
  MillerSched s;
  int step = 0;
  ZK_NOUNROLL for (;;) {
    if (s.phase == 0) proj_dbl_step(R, l, k);
    else proj_add_step(R, s.phase == 1 ? Q : s.phase == 2 ? Q1 : Q2, l);
    out[step].c0 = l.c0; out[step].c1 = l.c1; out[step].c3 = l.c3;
    step++;
    if (!s.advance()) break;
  }
}
// f = Miller value of (table, P) on the lane group; P at infinity -> 1
ZK_D void c12_miller(C12& f, const LineRec* __restrict__ lines, const G1P& P, const CoopLane& L) {
  c12_set_one(f, L);
  if (P.inf) return;               // uniform inside the group (all lanes hold the same P)
  MillerSched s;
  int step = 0;
  ZK_NOUNROLL for (;;) {
    if (s.is_dbl()) c12_sqr(f, f, L);
    const Fq2 c0 = fq2_scale(lines[step].c0, P.y), c1 = fq2_scale(lines[step].c1, P.x);
    c12_mul_line(f, c0, c1, lines[step].c3, L);
    step++;
    if (!s.advance()) break;
  }
}

// ------------------------------------------------------------------------------ G2 subgroup test
// The batched check adds up pairings under random weights, which is only sound for points of order r; the twist has a large
// cofactor, so B is tested with the endomorphism psi (untwist-Frobenius-twist): [x+1]P + psi([x]P) + psi^2([x]P) == psi^3([2x]P)
// (checked numerically against the oracle arithmetic in tests/dev/g2_subgroup_check.py).  A 63-bit scalar multiplication
// instead of a 254-bit one.
ZK_HD zk::G2Xyzz g2_psi(const zk::G2Xyzz& p, const PairingConsts& k) {
  // on XYZZ coordinates: x = X/ZZ, y = Y/ZZZ; conj is a field automorphism, so psi acts coordinate-wise with the constants on X, Y
  zk::G2Xyzz r;
  r.X = fq2_conj(p.X) * k.g12; r.Y = fq2_conj(p.Y) * k.g13; r.ZZ = fq2_conj(p.ZZ); r.ZZZ = fq2_conj(p.ZZZ);
  return r;
}
ZK_HD bool g2_in_subgroup(const G2P& Q, const PairingConsts& k) {
  if (Q.inf) return true;
  zk::G2Affine q; q.x = Q.x; q.y = Q.y;
  const uint64_t x = 0x44e992b44a6909f1ull;
  zk::G2Xyzz a = zk::G2Xyzz::from_affine(q);
  ZK_NOUNROLL for (int i = 61; i >= 0; i--) {
    a = zk::xyzz_dbl(a);
    if ((x >> i) & 1) zk::xyzz_madd(a, q, false);
  }
  zk::G2Xyzz b = g2_psi(a, k);
  zk::xyzz_madd(a, q, false);
  zk::G2Xyzz res = g2_psi(b, k);
  zk::G2Xyzz c = res;
  zk::xyzz_add(c, b);
  zk::xyzz_add(c, a);
  res = g2_psi(res, k);
  res = zk::xyzz_dbl(res);
  zk::xyzz_add(res, zk::xyzz_neg(c));
  return res.is_inf();
}
}  // namespace zkp
