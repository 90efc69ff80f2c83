// BN254 arithmetic for the sm_100a kernels: Fq / Fr in 8 x 32-bit-limb Montgomery form
// (R = 2^256, the representation the snarkjs `.zkey` stores its points in -- SURVEY A.5),
// Fq2 = Fq[u]/(u^2+1), and the short-Weierstrass group law in XYZZ coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ = 0) templated over the coordinate field.
// The integer pipe (IMAD) is the bound for everything in here; no tensor-core shape exists.
#pragma once
#include "zkfl_rt.h"

namespace zk {

// ------------------------------------------------------------------------------ parameters
struct FqP {
  static ZK_HD uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  static ZK_HD uint32_t r2(int i) {
    constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return m[i];
  }
  static ZK_HD uint32_t one(int i) {  // R mod q
    constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  static ZK_HD uint32_t inv() { return 0xe4866389u; }  // -q^-1 mod 2^32
};
struct FrP {
  static ZK_HD uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  static ZK_HD uint32_t r2(int i) {
    constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return m[i];
  }
  static ZK_HD uint32_t one(int i) {  // R mod r
    constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  static ZK_HD uint32_t inv() { return 0xefffffffu; }
};

// ------------------------------------------------------------------------------ PTX carry-chain helpers (device)
#if defined(__CUDA_ARCH__) && !defined(ZKFL_PORTABLE_MUL)
#define ZKFL_PTX_MUL 1
namespace ptx {
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
}  // namespace ptx
#endif

// ------------------------------------------------------------------------------ prime field
template <class P>
struct alignas(16) Fp {
  uint32_t v[8];

  static ZK_HD Fp zero() { Fp r; ZK_UNROLL for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
  static ZK_HD Fp one() { Fp r; ZK_UNROLL for (int i = 0; i < 8; i++) r.v[i] = P::one(i); return r; }
  static ZK_HD Fp r2() { Fp r; ZK_UNROLL for (int i = 0; i < 8; i++) r.v[i] = P::r2(i); return r; }
  ZK_HD bool is_zero() const { uint32_t o = 0; ZK_UNROLL for (int i = 0; i < 8; i++) o |= v[i]; return o == 0; }
  ZK_HD bool operator==(const Fp& b) const { uint32_t o = 0; ZK_UNROLL for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i]; return o == 0; }

  // r = t - mod if t >= mod else t   (t < 2*mod)
  static ZK_HD Fp reduce_once(const uint32_t* t) {
    uint32_t s[8]; uint64_t bw = 0;
    ZK_UNROLL for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)t[i] - P::mod(i) - bw; s[i] = (uint32_t)d; bw = (d >> 32) & 1; }
    Fp r; uint32_t keep = (uint32_t)0 - (uint32_t)bw;  // all ones when t < mod
    ZK_UNROLL for (int i = 0; i < 8; i++) r.v[i] = (t[i] & keep) | (s[i] & ~keep);
    return r;
  }
  friend ZK_HD Fp operator+(const Fp& a, const Fp& b) {
#ifdef ZKFL_PTX_MUL
    // carry chains in PTX: 8 adds, 9 subtracts, 8 selects (the portable form below costs about twice the instructions)
    uint32_t t[8], u[8];
    t[0] = ptx::add_cc(a.v[0], b.v[0]);
    ZK_UNROLL for (int i = 1; i < 7; i++) t[i] = ptx::addc_cc(a.v[i], b.v[i]);
    t[7] = ptx::addc(a.v[7], b.v[7]);                       // a + b < 2^255: no carry out
    u[0] = ptx::sub_cc(t[0], P::mod(0));
    ZK_UNROLL for (int i = 1; i < 8; i++) u[i] = ptx::subc_cc(t[i], P::mod(i));
    const uint32_t borrow = ptx::subc(0, 0);                // all ones when t < p
    Fp r;
    ZK_UNROLL for (int i = 0; i < 8; i++) r.v[i] = borrow ? t[i] : u[i];
    return r;
#else
    uint32_t t[8]; uint64_t c = 0;
    ZK_UNROLL for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; t[i] = (uint32_t)c; c >>= 32; }
    return reduce_once(t);  // a + b < 2^255: no carry out
#endif
  }
  friend ZK_HD Fp operator-(const Fp& a, const Fp& b) {
#ifdef ZKFL_PTX_MUL
    uint32_t t[8];
    t[0] = ptx::sub_cc(a.v[0], b.v[0]);
    ZK_UNROLL for (int i = 1; i < 8; i++) t[i] = ptx::subc_cc(a.v[i], b.v[i]);
    const uint32_t borrow = ptx::subc(0, 0);                // all ones when a < b: add p back
    Fp r;
    r.v[0] = ptx::add_cc(t[0], P::mod(0) & borrow);
    ZK_UNROLL for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(t[i], P::mod(i) & borrow);
    r.v[7] = ptx::addc(t[7], P::mod(7) & borrow);
    return r;
#else
    uint32_t t[8]; uint64_t bw = 0;
    ZK_UNROLL for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a.v[i] - b.v[i] - bw; t[i] = (uint32_t)d; bw = (d >> 32) & 1; }
    uint32_t msk = (uint32_t)0 - (uint32_t)bw; uint64_t c = 0; Fp r;
    ZK_UNROLL for (int i = 0; i < 8; i++) { c += (uint64_t)t[i] + (P::mod(i) & msk); r.v[i] = (uint32_t)c; c >>= 32; }
    return r;
#endif
  }
  ZK_HD Fp neg() const { return zero() - *this; }
  ZK_HD Fp dbl() const { return *this + *this; }

  // Montgomery product a*b/R, CIOS; p < 2^254 keeps the running value in 9 words.
  // `mul_inline` is the body; `operator*` is a by-value LEAF CALL on the device (operands and result
  // travel in registers, no stack frame), which keeps the many group-law call sites compact.  The hot
  // bucket-accumulation path uses mul_inline directly so ptxas can schedule across products.
  static ZK_HD Fp mul_inline(const Fp& a, const Fp& b) {
#ifdef ZKFL_PTX_MUL
    // Even/odd paired carry chains: every 32x32->64 product is a (mad.lo.cc, madc.hi.cc) pair on two ADJACENT
    // accumulator words, which ptxas fuses into one IMAD.WIDE.U32(.X) with the carry in a predicate. Products of
    // the even limbs accumulate in ev[] (word k has weight 2^(32k)), products of the odd limbs in od[] (word k has
    // weight 2^(32(k+1))); after each Montgomery step the arrays swap roles (division by 2^32).  128 wide MACs +
    // 8 quotient products per modular product; measured IMAD.WIDE rate on B200: ~35 per clock per SM.
    uint32_t ev[8], od[8];
    ZK_UNROLL for (int i = 0; i < 8; i++) {
      uint32_t* X = (i & 1) ? od : ev;   // plays "even" in this step
      uint32_t* Y = (i & 1) ? ev : od;   // plays "odd"
      const uint32_t bi = b.v[i];
      if (i == 0) {
        ZK_UNROLL for (int j = 0; j < 8; j += 2) {   // 64-bit products: one IMAD.WIDE each instead of IMAD + half-rate IMAD.HI
          const uint64_t py = (uint64_t)a.v[j + 1] * bi, px = (uint64_t)a.v[j] * bi;
          Y[j] = (uint32_t)py; Y[j + 1] = (uint32_t)(py >> 32);
          X[j] = (uint32_t)px; X[j + 1] = (uint32_t)(px >> 32);
        }
      } else {
        // previous total / 2^32: Y[0] is zero, Y[1] joins X[0], Y shifts down two words while taking the odd products
        X[0] = ptx::add_cc(X[0], Y[1]);
        ZK_UNROLL for (int j = 0; j < 6; j += 2) {
          Y[j] = ptx::madc_lo_cc(a.v[j + 1], bi, Y[j + 2]);
          Y[j + 1] = ptx::madc_hi_cc(a.v[j + 1], bi, Y[j + 3]);
        }
        Y[6] = ptx::madc_lo_cc(a.v[7], bi, 0);
        Y[7] = ptx::madc_hi(a.v[7], bi, 0);
        X[0] = ptx::mad_lo_cc(a.v[0], bi, X[0]); X[1] = ptx::madc_hi_cc(a.v[0], bi, X[1]);
        ZK_UNROLL for (int j = 2; j < 8; j += 2) {
          X[j] = ptx::madc_lo_cc(a.v[j], bi, X[j]);
          X[j + 1] = ptx::madc_hi_cc(a.v[j], bi, X[j + 1]);
        }
        Y[7] = ptx::addc(Y[7], 0);
      }
      const uint32_t m = X[0] * P::inv();
      Y[0] = ptx::mad_lo_cc(m, P::mod(1), Y[0]); Y[1] = ptx::madc_hi_cc(m, P::mod(1), Y[1]);
      ZK_UNROLL for (int j = 2; j < 8; j += 2) {
        Y[j] = ptx::madc_lo_cc(m, P::mod(j + 1), Y[j]);
        Y[j + 1] = ptx::madc_hi_cc(m, P::mod(j + 1), Y[j + 1]);
      }
      X[0] = ptx::mad_lo_cc(m, P::mod(0), X[0]); X[1] = ptx::madc_hi_cc(m, P::mod(0), X[1]);
      ZK_UNROLL for (int j = 2; j < 8; j += 2) {
        X[j] = ptx::madc_lo_cc(m, P::mod(j), X[j]);
        X[j + 1] = ptx::madc_hi_cc(m, P::mod(j), X[j + 1]);
      }
      Y[7] = ptx::addc(Y[7], 0);
    }
    // after 8 steps od[] played "even" last (od[0] == 0): result = ev + od / 2^32, below 2p
    uint32_t t0 = ptx::add_cc(ev[0], od[1]), t1 = ptx::addc_cc(ev[1], od[2]), t2 = ptx::addc_cc(ev[2], od[3]),
             t3 = ptx::addc_cc(ev[3], od[4]), t4 = ptx::addc_cc(ev[4], od[5]), t5 = ptx::addc_cc(ev[5], od[6]),
             t6 = ptx::addc_cc(ev[6], od[7]), t7 = ptx::addc(ev[7], 0);
    uint32_t s0 = ptx::sub_cc(t0, P::mod(0)), s1 = ptx::subc_cc(t1, P::mod(1)), s2 = ptx::subc_cc(t2, P::mod(2)),
             s3 = ptx::subc_cc(t3, P::mod(3)), s4 = ptx::subc_cc(t4, P::mod(4)), s5 = ptx::subc_cc(t5, P::mod(5)),
             s6 = ptx::subc_cc(t6, P::mod(6)), s7 = ptx::subc_cc(t7, P::mod(7));
    const uint32_t borrow = ptx::subc(0, 0);  // 0xffffffff when t < p
    Fp r;
    r.v[0] = borrow ? t0 : s0; r.v[1] = borrow ? t1 : s1; r.v[2] = borrow ? t2 : s2; r.v[3] = borrow ? t3 : s3;
    r.v[4] = borrow ? t4 : s4; r.v[5] = borrow ? t5 : s5; r.v[6] = borrow ? t6 : s6; r.v[7] = borrow ? t7 : s7;
    return r;
#else
    uint32_t t[9];
    ZK_UNROLL for (int i = 0; i < 9; i++) t[i] = 0;
    ZK_UNROLL for (int i = 0; i < 8; i++) {
      uint64_t c = 0;
      const uint32_t bi = b.v[i];
      ZK_UNROLL for (int j = 0; j < 8; j++) { c += (uint64_t)a.v[j] * bi + t[j]; t[j] = (uint32_t)c; c >>= 32; }
      c += t[8]; t[8] = (uint32_t)c;
      const uint32_t m = t[0] * P::inv();
      c = (uint64_t)m * P::mod(0) + t[0]; c >>= 32;
      ZK_UNROLL for (int j = 1; j < 8; j++) { c += (uint64_t)m * P::mod(j) + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
      c += t[8]; t[7] = (uint32_t)c; t[8] = (uint32_t)(c >> 32);
    }
    return reduce_once(t);
#endif
  }
  static ZK_HD_NOINLINE Fp mul_call(Fp a, Fp b) { return mul_inline(a, b); }
  friend ZK_HD Fp operator*(const Fp& a, const Fp& b) { return mul_call(a, b); }
  // hot-path product of the bucket accumulation. Measured on B200 (1024 proofs per step): inlined into the G1 mixed add it
  // removes ~370 call-ABI moves per addition (167 -> 152 ms); inside the Fq2 product (G2 mixed add, 28 products) the
  // inlined body overflows the instruction cache (92 -> 112 ms), so Fq2 keeps the leaf call (see Fq2::mul_hot).
  static ZK_HD Fp mul_hot(const Fp& a, const Fp& b) { return mul_inline(a, b); }
  static ZK_HD Fp sqr_hot(const Fp& a) { return mul_hot(a, a); }

  // ------------------------------------------------------------------ lazy reduction building blocks
  // 512-bit unreduced product and a separate Montgomery reduction (limb-exact model: tests/dev/wide_model.py).
  // For sharing ONE reduction between several products: Fq2 Karatsuba (3 wide products + 2 reductions = 336 wide MACs instead
  // of 408) and Y3 = R*(Q - X3) - Y1*PPP in the G1 mixed add (200 instead of 272).  MEASURED SLOWER on B200 (G2 bucket
  // accumulation 92 -> 113 ms at 1024 proofs): the 512-bit add/sub glue and the explicit shifts of a separate reduction cost
  // more issue slots than the 72 wide MACs saved, so it is opt-in (-DZKFL_LAZY_REDUCTION) and off by default.
  struct Wide { uint32_t v[16]; };

  // a + b without reduction (inputs < p, result < 2p < 2^255)
  static ZK_HD Fp add_noreduce(const Fp& a, const Fp& b) {
    Fp r; uint64_t c = 0;
    ZK_UNROLL for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    return r;
  }
  static ZK_HD Wide wide_add(const Wide& x, const Wide& y) {   // no overflow by the callers' bounds
    Wide r; uint64_t c = 0;
    ZK_UNROLL for (int i = 0; i < 16; i++) { c += (uint64_t)x.v[i] + y.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    return r;
  }
  static ZK_HD Wide wide_sub(const Wide& x, const Wide& y) {   // x >= y by the callers' bounds
    Wide r; uint64_t bw = 0;
    ZK_UNROLL for (int i = 0; i < 16; i++) { uint64_t d = (uint64_t)x.v[i] - y.v[i] - bw; r.v[i] = (uint32_t)d; bw = (d >> 32) & 1; }
    return r;
  }
  static ZK_HD Wide wide_add_pR(const Wide& x) {               // x + p * 2^256
    Wide r = x; uint64_t c = 0;
    ZK_UNROLL for (int i = 0; i < 8; i++) { c += (uint64_t)x.v[8 + i] + P::mod(i); r.v[8 + i] = (uint32_t)c; c >>= 32; }
    return r;
  }

  static ZK_HD Wide mul_wide(const Fp& a, const Fp& b) {
    Wide r;
#ifdef ZKFL_PTX_MUL
    // even/odd paired chains as in mul_inline, without the reduction rows. E word k has weight 2^(32k), O word k has
    // weight 2^(32(k+1)); each chain's carry lands in a word that so far only holds carries.
    uint32_t E[17], O[17];
    ZK_UNROLL for (int k = 0; k < 17; k++) { E[k] = 0; O[k] = 0; }
    ZK_UNROLL for (int i = 0; i < 8; i++) {
      const uint32_t bi = b.v[i];
      const int pe = i & 1;            // parity of the a-limbs whose products are word-aligned with E in this row
      const int e0 = i + pe;           // first E word of the row
      E[e0] = ptx::mad_lo_cc(a.v[pe], bi, E[e0]); E[e0 + 1] = ptx::madc_hi_cc(a.v[pe], bi, E[e0 + 1]);
      ZK_UNROLL for (int t = 1; t < 4; t++) {
        E[e0 + 2 * t] = ptx::madc_lo_cc(a.v[pe + 2 * t], bi, E[e0 + 2 * t]);
        E[e0 + 2 * t + 1] = ptx::madc_hi_cc(a.v[pe + 2 * t], bi, E[e0 + 2 * t + 1]);
      }
      E[e0 + 8] = ptx::addc(E[e0 + 8], 0);
      const int po = 1 - pe;           // the other parity goes to O
      const int o0 = i + po - 1;       // first O word of the row (O is offset by one word)
      O[o0] = ptx::mad_lo_cc(a.v[po], bi, O[o0]); O[o0 + 1] = ptx::madc_hi_cc(a.v[po], bi, O[o0 + 1]);
      ZK_UNROLL for (int t = 1; t < 4; t++) {
        O[o0 + 2 * t] = ptx::madc_lo_cc(a.v[po + 2 * t], bi, O[o0 + 2 * t]);
        O[o0 + 2 * t + 1] = ptx::madc_hi_cc(a.v[po + 2 * t], bi, O[o0 + 2 * t + 1]);
      }
      O[o0 + 8] = ptx::addc(O[o0 + 8], 0);
    }
    r.v[0] = E[0];
    r.v[1] = ptx::add_cc(E[1], O[0]);
    ZK_UNROLL for (int k = 2; k < 15; k++) r.v[k] = ptx::addc_cc(E[k], O[k - 1]);
    r.v[15] = ptx::addc(E[15], O[14]);
#else
    uint32_t t[16];
    ZK_UNROLL for (int i = 0; i < 16; i++) t[i] = 0;
    ZK_UNROLL for (int i = 0; i < 8; i++) {
      uint64_t c = 0;
      ZK_UNROLL for (int j = 0; j < 8; j++) { c += (uint64_t)a.v[j] * b.v[i] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
      t[i + 8] = (uint32_t)c;
    }
    ZK_UNROLL for (int i = 0; i < 16; i++) r.v[i] = t[i];
#endif
    return r;
  }

  // t * 2^-256 mod p for t < 2 * p * 2^256
  static ZK_HD Fp redc(const Wide& t) {
    uint32_t r[9];
#ifdef ZKFL_PTX_MUL
    uint32_t X[8], Y[8];
    ZK_UNROLL for (int k = 0; k < 8; k++) { X[k] = t.v[k]; Y[k] = 0; }
    ZK_UNROLL for (int i = 0; i < 8; i++) {
      const uint32_t m = X[0] * P::inv();
      Y[0] = ptx::mad_lo_cc(m, P::mod(1), Y[0]); Y[1] = ptx::madc_hi_cc(m, P::mod(1), Y[1]);
      ZK_UNROLL for (int j = 2; j < 8; j += 2) { Y[j] = ptx::madc_lo_cc(m, P::mod(j + 1), Y[j]); Y[j + 1] = ptx::madc_hi_cc(m, P::mod(j + 1), Y[j + 1]); }
      X[0] = ptx::mad_lo_cc(m, P::mod(0), X[0]); X[1] = ptx::madc_hi_cc(m, P::mod(0), X[1]);
      ZK_UNROLL for (int j = 2; j < 8; j += 2) { X[j] = ptx::madc_lo_cc(m, P::mod(j), X[j]); X[j + 1] = ptx::madc_hi_cc(m, P::mod(j), X[j + 1]); }
      Y[7] = ptx::addc(Y[7], 0);
      // divide by 2^32 (X[0] is zero): X' = Y with X'[0] += X[1]; Y' = X[2..7], carry rippling; then bring in t[8 + i]
      uint32_t nx[8], ny[8];
      nx[0] = ptx::add_cc(Y[0], X[1]);
      ZK_UNROLL for (int k = 0; k < 6; k++) ny[k] = ptx::addc_cc(X[k + 2], 0);
      ny[6] = ptx::addc(0, 0);
      ZK_UNROLL for (int k = 1; k < 7; k++) nx[k] = Y[k];
      nx[7] = ptx::add_cc(Y[7], t.v[8 + i]);
      ny[7] = ptx::addc(0, 0);
      ZK_UNROLL for (int k = 0; k < 8; k++) { X[k] = nx[k]; Y[k] = ny[k]; }
    }
    r[0] = X[0];
    r[1] = ptx::add_cc(X[1], Y[0]);
    ZK_UNROLL for (int k = 2; k < 8; k++) r[k] = ptx::addc_cc(X[k], Y[k - 1]);
    r[8] = ptx::addc(0, Y[7]);
#else
    uint32_t w[17];
    ZK_UNROLL for (int i = 0; i < 16; i++) w[i] = t.v[i];
    w[16] = 0;
    ZK_UNROLL for (int i = 0; i < 8; i++) {
      const uint32_t m = w[i] * P::inv();
      uint64_t c = 0;
      ZK_UNROLL for (int j = 0; j < 8; j++) { c += (uint64_t)m * P::mod(j) + w[i + j]; w[i + j] = (uint32_t)c; c >>= 32; }
      for (int k = i + 8; k < 17 && c; k++) { c += w[k]; w[k] = (uint32_t)c; c >>= 32; }
    }
    ZK_UNROLL for (int i = 0; i < 9; i++) r[i] = w[8 + i];
#endif
    // value < 3p: at most two subtractions
    ZK_UNROLL for (int round = 0; round < 2; round++) {
      uint32_t d[9]; uint64_t bw = 0;
      ZK_UNROLL for (int i = 0; i < 9; i++) { uint64_t q = (uint64_t)r[i] - (i < 8 ? P::mod(i) : 0u) - bw; d[i] = (uint32_t)q; bw = (q >> 32) & 1; }
      const uint32_t keep = (uint32_t)0 - (uint32_t)bw;   // all ones when r < p
      ZK_UNROLL for (int i = 0; i < 9; i++) r[i] = (r[i] & keep) | (d[i] & ~keep);
    }
    Fp o; ZK_UNROLL for (int i = 0; i < 8; i++) o.v[i] = r[i];
    return o;
  }
  // (a*b - c*d) / R mod p with one reduction
  static ZK_HD_NOINLINE Fp diff_of_products_call(Fp a, Fp b, Fp c, Fp d) {
    return redc(wide_sub(wide_add_pR(mul_wide(a, b)), mul_wide(c, d)));
  }
  static ZK_HD Fp diff_of_products(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
#ifndef ZKFL_LAZY_REDUCTION
    return mul_hot(a, b) - mul_hot(c, d);
#else
    return diff_of_products_call(a, b, c, d);
#endif
  }
  ZK_HD Fp sqr() const { return *this * *this; }
  ZK_HD Fp to_mont() const { return *this * r2(); }
  ZK_HD Fp from_mont() const { Fp o = zero(); o.v[0] = 1; return *this * o; }

  // Binary extended Euclid (HAC 14.61): ~380 halvings + ~180 subtractions of 256-bit integers, about a fifth of the issue
  // slots of the Fermat ladder below.  Used for the ONE shared inversion of a batch-affine round (k_msm_accumulate_affine),
  // where every lane of the warp inverts the same value, so the data-dependent branches are warp-uniform.
  // Input a*R (Montgomery), output a^-1 * R; 0 -> 0.
  ZK_HD Fp inv_gcd() const {
    if (is_zero()) return *this;
    uint32_t u[8], w[8], x1[8], x2[8];
    ZK_UNROLL for (int i = 0; i < 8; i++) { u[i] = v[i]; w[i] = P::mod(i); x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    // invariant: x1 * (aR) = u, x2 * (aR) = w  (mod p); all four stay below p
    auto is_one = [](const uint32_t* a) { uint32_t o = a[0] ^ 1u; ZK_UNROLL for (int i = 1; i < 8; i++) o |= a[i]; return o == 0; };
    auto shr1 = [](uint32_t* a, uint32_t top) {
      ZK_UNROLL for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
      a[7] = (a[7] >> 1) | (top << 31);
    };
    auto halve_mod = [&](uint32_t* a) {   // a / 2 mod p
      uint32_t msk = (uint32_t)0 - (a[0] & 1u); uint64_t c = 0;
      ZK_UNROLL for (int i = 0; i < 8; i++) { c += (uint64_t)a[i] + (P::mod(i) & msk); a[i] = (uint32_t)c; c >>= 32; }
      shr1(a, (uint32_t)c);
    };
    auto sub_to = [](uint32_t* a, const uint32_t* b) -> uint32_t {   // a -= b, returns the borrow
      uint64_t bw = 0;
      ZK_UNROLL for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a[i] - b[i] - bw; a[i] = (uint32_t)d; bw = (d >> 32) & 1; }
      return (uint32_t)bw;
    };
    auto sub_mod = [&](uint32_t* a, const uint32_t* b) {              // a = a - b mod p
      uint32_t msk = (uint32_t)0 - sub_to(a, b); uint64_t c = 0;
      ZK_UNROLL for (int i = 0; i < 8; i++) { c += (uint64_t)a[i] + (P::mod(i) & msk); a[i] = (uint32_t)c; c >>= 32; }
    };
    auto geq = [](const uint32_t* a, const uint32_t* b) {
      uint64_t bw = 0;
      ZK_UNROLL for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a[i] - b[i] - bw; bw = (d >> 32) & 1; }
      return bw == 0;
    };
    ZK_NOUNROLL while (!is_one(u) && !is_one(w)) {
      ZK_NOUNROLL while (!(u[0] & 1u)) { shr1(u, 0); halve_mod(x1); }
      ZK_NOUNROLL while (!(w[0] & 1u)) { shr1(w, 0); halve_mod(x2); }
      if (geq(u, w)) { sub_to(u, w); sub_mod(x1, x2); } else { sub_to(w, u); sub_mod(x2, x1); }
    }
    Fp r; const bool from_u = is_one(u);
    ZK_UNROLL for (int i = 0; i < 8; i++) r.v[i] = from_u ? x1[i] : x2[i];
    // r = (aR)^-1; wanted a^-1 * R = r * R^2 = montmul(r, R^3)
    return r * (r2() * r2());
  }
  ZK_HD Fp inv() const {  // Fermat, exponent p - 2 (setup / affine conversion only)
    Fp r = one();
    ZK_NOUNROLL for (int i = 253; i >= 0; i--) {
      r = r.sqr();
      uint32_t w = P::mod(i >> 5);
      if ((i >> 5) == 0) w -= 2;  // mod(0) is odd and >= 3: no borrow
      if ((w >> (i & 31)) & 1) r = r * *this;
    }
    return r;
  }
};
typedef Fp<FqP> Fq;
typedef Fp<FrP> Fr;

// ------------------------------------------------------------------------------ Fq2
struct alignas(16) Fq2 {
  Fq a, b;  // a + b*u
  static ZK_HD Fq2 zero() { Fq2 r; r.a = Fq::zero(); r.b = Fq::zero(); return r; }
  static ZK_HD Fq2 one() { Fq2 r; r.a = Fq::one(); r.b = Fq::zero(); return r; }
  ZK_HD bool is_zero() const { return a.is_zero() && b.is_zero(); }
  ZK_HD bool operator==(const Fq2& o) const { return a == o.a && b == o.b; }
  friend ZK_HD Fq2 operator+(const Fq2& x, const Fq2& y) { Fq2 r; r.a = x.a + y.a; r.b = x.b + y.b; return r; }
  friend ZK_HD Fq2 operator-(const Fq2& x, const Fq2& y) { Fq2 r; r.a = x.a - y.a; r.b = x.b - y.b; return r; }
  ZK_HD Fq2 neg() const { Fq2 r; r.a = a.neg(); r.b = b.neg(); return r; }
  ZK_HD Fq2 dbl() const { Fq2 r; r.a = a.dbl(); r.b = b.dbl(); return r; }
  // ZKFL_FQ2_OPS_AS_ONE_CALL (set by verify.cu): the generic product / square are the one-call forms below, whose three (two)
  // inlined Fq products interleave -- the verifier's kernels are serial chains on nearly idle SMs, where three back-to-back
  // Fq calls cost three full product latencies (~1 500 cycles each)
  friend ZK_HD Fq2 operator*(const Fq2& x, const Fq2& y) {
#if defined(ZKFL_FQ2_OPS_AS_ONE_CALL) && defined(__CUDA_ARCH__)   // device code of that unit only
    return mul_call2(x, y);
#else
    Fq aa = x.a * y.a, bb = x.b * y.b, s = (x.a + x.b) * (y.a + y.b);
    Fq2 r; r.a = aa - bb; r.b = s - aa - bb; return r;
#endif
  }
  ZK_HD Fq2 sqr() const {
#if defined(ZKFL_FQ2_OPS_AS_ONE_CALL) && defined(__CUDA_ARCH__)   // device code of that unit only
    return sqr_call2(*this);
#else
    Fq t = a * b; Fq2 r; r.a = (a + b) * (a - b); r.b = t.dbl(); return r;
#endif
  }
  static ZK_HD Fq2 sqr_hot(const Fq2& x) {   // (a + b)(a - b) + 2ab u: two products instead of three
#if !defined(ZKFL_FQ2_PER_FQ_CALLS) && !defined(ZKFL_LAZY_REDUCTION)
    if (true) return sqr_call2(x);
#endif
    Fq t = Fq::mul_call(x.a, x.b);
    Fq2 r; r.a = Fq::mul_call(x.a + x.b, x.a - x.b); r.b = t.dbl(); return r;
  }
#ifdef ZKFL_LAZY_REDUCTION
  // Karatsuba with lazy reduction: three 512-bit products, two Montgomery reductions
  static ZK_HD_NOINLINE Fq2 mul_lazy_call(Fq2 x, Fq2 y) {
    Fq::Wide aa = Fq::mul_wide(x.a, y.a), bb = Fq::mul_wide(x.b, y.b);
    Fq::Wide ss = Fq::mul_wide(Fq::add_noreduce(x.a, x.b), Fq::add_noreduce(y.a, y.b));
    Fq2 r;
    r.a = Fq::redc(Fq::wide_sub(Fq::wide_add_pR(aa), bb));
    r.b = Fq::redc(Fq::wide_sub(Fq::wide_sub(ss, aa), bb));
    return r;
  }
#endif
  // whole Fq2 product / square as ONE register-passed leaf call holding three (two) inlined Fq products: a third of the
  // call traffic of per-Fq calls, and ptxas can interleave the independent carry chains inside
  static ZK_HD_NOINLINE Fq2 mul_call2(Fq2 x, Fq2 y) {
    Fq aa = Fq::mul_inline(x.a, y.a), bb = Fq::mul_inline(x.b, y.b), s = Fq::mul_inline(x.a + x.b, y.a + y.b);
    Fq2 r; r.a = aa - bb; r.b = s - aa - bb; return r;
  }
  static ZK_HD_NOINLINE Fq2 sqr_call2(Fq2 x) {
    Fq t = Fq::mul_inline(x.a, x.b);
    Fq2 r; r.a = Fq::mul_inline(x.a + x.b, x.a - x.b); r.b = t.dbl(); return r;
  }
  // default: measured 82.5 -> 80.1 ms per 1024 proofs for the G2 accumulation against three per-Fq calls (-DZKFL_FQ2_PER_FQ_CALLS)
  static ZK_HD Fq2 mul_hot(const Fq2& x, const Fq2& y) {
#if !defined(ZKFL_FQ2_PER_FQ_CALLS) && !defined(ZKFL_LAZY_REDUCTION)
    return mul_call2(x, y);
#elif !defined(ZKFL_LAZY_REDUCTION)
    Fq aa = Fq::mul_call(x.a, y.a), bb = Fq::mul_call(x.b, y.b), s = Fq::mul_call(x.a + x.b, y.a + y.b);
    Fq2 r; r.a = aa - bb; r.b = s - aa - bb; return r;
#else
    return mul_lazy_call(x, y);
#endif
  }
  static ZK_HD Fq2 diff_of_products(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) { return mul_hot(a, b) - mul_hot(c, d); }
  ZK_HD Fq2 inv() const { Fq d = (a.sqr() + b.sqr()).inv(); Fq2 r; r.a = a * d; r.b = (b * d).neg(); return r; }
  ZK_HD Fq2 inv_gcd() const { Fq d = (a.sqr() + b.sqr()).inv_gcd(); Fq2 r; r.a = a * d; r.b = (b * d).neg(); return r; }
  ZK_HD Fq2 from_mont() const { Fq2 r; r.a = a.from_mont(); r.b = b.from_mont(); return r; }
};

// ------------------------------------------------------------------------------ curve points
template <class F> struct alignas(16) Affine {
  F x, y;  // Montgomery coordinates; (0,0) encodes infinity (zkey convention)
  ZK_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
};
template <class F> struct alignas(16) Xyzz {
  F X, Y, ZZ, ZZZ;
  static ZK_HD Xyzz infinity() { Xyzz r; r.X = F::zero(); r.Y = F::zero(); r.ZZ = F::zero(); r.ZZZ = F::zero(); return r; }
  ZK_HD bool is_inf() const { return ZZ.is_zero(); }
  static ZK_HD Xyzz from_affine(const Affine<F>& p) {
    if (p.is_inf()) return infinity();
    Xyzz r; r.X = p.x; r.Y = p.y; r.ZZ = F::one(); r.ZZZ = F::one(); return r;
  }
};

template <class F> ZK_HD Xyzz<F> xyzz_dbl(const Xyzz<F>& p) {  // dbl-2008-s-1, a = 0
  if (p.is_inf()) return p;
  F U = p.Y.dbl(), V = U.sqr(), W = U * V, S = p.X * V;
  F M = p.X.sqr(); M = M.dbl() + M;
  Xyzz<F> r;
  r.X = M.sqr() - S.dbl();
  r.Y = M * (S - r.X) - W * p.Y;
  r.ZZ = V * p.ZZ;
  r.ZZZ = W * p.ZZZ;
  return r;
}
template <class F> ZK_HD Xyzz<F> xyzz_dbl_affine(const Affine<F>& p) {  // mdbl-2008-s-1
  F U = p.y.dbl(), V = U.sqr(), W = U * V, S = p.x * V;
  F M = p.x.sqr(); M = M.dbl() + M;
  Xyzz<F> r;
  r.X = M.sqr() - S.dbl();
  r.Y = M * (S - r.X) - W * p.y;
  r.ZZ = V;
  r.ZZZ = W;
  return r;
}
// acc += q (affine; infinity bases are skipped); `negate` flips q first.  Hot path of the MSM bucket
// accumulation (products go through mul_hot: inlined for Fq, a leaf call inside the Fq2 product -- see Fp::mul_hot).
template <class F> ZK_HD void xyzz_madd(Xyzz<F>& acc, const Affine<F>& q0, bool negate) {  // madd-2008-s
  if (q0.is_inf()) return;
  Affine<F> q = q0;
  if (negate) q.y = q.y.neg();
  if (acc.is_inf()) { acc = Xyzz<F>::from_affine(q); return; }
  F U2 = F::mul_hot(q.x, acc.ZZ), S2 = F::mul_hot(q.y, acc.ZZZ);
  F Pp = U2 - acc.X, Rr = S2 - acc.Y;
  if (Pp.is_zero()) {
    if (Rr.is_zero()) acc = xyzz_dbl_affine(q); else acc = Xyzz<F>::infinity();
    return;
  }
  F PP = F::sqr_hot(Pp), PPP = F::mul_hot(Pp, PP), Qq = F::mul_hot(acc.X, PP);
  F X3 = F::sqr_hot(Rr) - PPP - Qq.dbl();
  acc.Y = F::diff_of_products(Rr, Qq - X3, acc.Y, PPP);
  acc.X = X3;
  acc.ZZ = F::mul_hot(acc.ZZ, PP);
  acc.ZZZ = F::mul_hot(acc.ZZZ, PPP);
}
template <class F> ZK_HD void xyzz_add(Xyzz<F>& acc, const Xyzz<F>& q) {  // add-2008-s
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  F U1 = acc.X * q.ZZ, U2 = q.X * acc.ZZ, S1 = acc.Y * q.ZZZ, S2 = q.Y * acc.ZZZ;
  F Pp = U2 - U1, Rr = S2 - S1;
  if (Pp.is_zero()) {
    if (Rr.is_zero()) acc = xyzz_dbl(acc); else acc = Xyzz<F>::infinity();
    return;
  }
  F PP = Pp.sqr(), PPP = Pp * PP, Qq = U1 * PP;
  F X3 = Rr.sqr() - PPP - Qq.dbl();
  acc.Y = Rr * (Qq - X3) - S1 * PPP;
  acc.X = X3;
  acc.ZZ = acc.ZZ * q.ZZ * PP;
  acc.ZZZ = acc.ZZZ * q.ZZZ * PPP;
}
// the same addition / doubling with the products of the hot path (Fq: inlined, Fq2: one call per Fq2 product with its three Fq
// products interleaved): for the FEW-ROW kernels (single proofs, split proofs, one MSM), which are chains of dependent additions on
// a nearly idle GPU -- with by-value leaf calls the 14 products of an addition run back to back (~1 500 cycles each), inlined the
// independent ones (U1, U2, S1, S2; PPP, Q; the two halves of Y3; ZZ, ZZZ) overlap.  One call site per kernel (2.6 k instructions).
template <class F> ZK_HD void xyzz_add_hot(Xyzz<F>& acc, const Xyzz<F>& q) {  // add-2008-s
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  F U1 = F::mul_hot(acc.X, q.ZZ), U2 = F::mul_hot(q.X, acc.ZZ), S1 = F::mul_hot(acc.Y, q.ZZZ), S2 = F::mul_hot(q.Y, acc.ZZZ);
  F ZZq = F::mul_hot(acc.ZZ, q.ZZ), ZZZq = F::mul_hot(acc.ZZZ, q.ZZZ);
  F Pp = U2 - U1, Rr = S2 - S1;
  if (Pp.is_zero()) {
    if (Rr.is_zero()) acc = xyzz_dbl(acc); else acc = Xyzz<F>::infinity();
    return;
  }
  F PP = F::sqr_hot(Pp), PPP = F::mul_hot(Pp, PP), Qq = F::mul_hot(U1, PP);
  F X3 = F::sqr_hot(Rr) - PPP - Qq.dbl();
  acc.Y = F::mul_hot(Rr, Qq - X3) - F::mul_hot(S1, PPP);
  acc.X = X3;
  acc.ZZ = F::mul_hot(ZZq, PP);
  acc.ZZZ = F::mul_hot(ZZZq, PPP);
}
template <class F> ZK_HD Xyzz<F> xyzz_dbl_hot(const Xyzz<F>& p) {  // dbl-2008-s-1, a = 0
  if (p.is_inf()) return p;
  F U = p.Y.dbl(), V = F::sqr_hot(U), W = F::mul_hot(U, V), S = F::mul_hot(p.X, V);
  F M = F::sqr_hot(p.X); M = M.dbl() + M;
  Xyzz<F> r;
  r.X = F::sqr_hot(M) - S.dbl();
  r.Y = F::mul_hot(M, S - r.X) - F::mul_hot(W, p.Y);
  r.ZZ = F::mul_hot(V, p.ZZ);
  r.ZZZ = F::mul_hot(W, p.ZZZ);
  return r;
}
template <class F> ZK_HD Xyzz<F> xyzz_neg(const Xyzz<F>& p) { Xyzz<F> r = p; r.Y = p.Y.neg(); return r; }

// Montgomery-form affine; infinity -> (0,0)
template <class F> ZK_HD Affine<F> xyzz_to_affine(const Xyzz<F>& p) {
  Affine<F> r;
  if (p.is_inf()) { r.x = F::zero(); r.y = F::zero(); return r; }
  F i3 = p.ZZZ.inv();
  F i2 = p.ZZ.sqr() * i3.sqr();  // 1/ZZ = ZZ^2 / ZZZ^2
  r.x = p.X * i2;
  r.y = p.Y * i3;
  return r;
}
// k * p, k = 8 canonical little-endian words (< 2^254)
template <class F> ZK_HD Xyzz<F> xyzz_scalar_mul(const Xyzz<F>& p, const uint32_t* k) {
  Xyzz<F> r = Xyzz<F>::infinity();
  ZK_NOUNROLL for (int i = 253; i >= 0; i--) {
    r = xyzz_dbl(r);
    if ((k[i >> 5] >> (i & 31)) & 1) xyzz_add(r, p);
  }
  return r;
}

ZK_HD Fq to_mont_any(const Fq& x) { return x.to_mont(); }
ZK_HD Fq2 to_mont_any(const Fq2& x) { Fq2 r; r.a = x.a.to_mont(); r.b = x.b.to_mont(); return r; }

// ------------------------------------------------------------------------------ warp exchange of field elements (device)
#if !defined(ZKFL_EMUL)
enum { ZK_SHFL_UP = 0, ZK_SHFL_DOWN = 1, ZK_SHFL_IDX = 2 };
template <int MODE, class P> __device__ __forceinline__ Fp<P> warp_shfl(const Fp<P>& x, uint32_t arg) {
  Fp<P> r;
  ZK_UNROLL for (int i = 0; i < 8; i++)
    r.v[i] = MODE == ZK_SHFL_UP ? __shfl_up_sync(0xffffffffu, x.v[i], arg)
           : MODE == ZK_SHFL_DOWN ? __shfl_down_sync(0xffffffffu, x.v[i], arg) : __shfl_sync(0xffffffffu, x.v[i], (int)arg);
  return r;
}
template <int MODE> __device__ __forceinline__ Fq2 warp_shfl(const Fq2& x, uint32_t arg) {
  Fq2 r; r.a = warp_shfl<MODE>(x.a, arg); r.b = warp_shfl<MODE>(x.b, arg); return r;
}
#endif

typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef Xyzz<Fq> G1Xyzz;
typedef Xyzz<Fq2> G2Xyzz;

}  // namespace zk
