// ORACLE (test infrastructure + CPU baseline; never linked into or called by the product).
//
// Fast multi-threaded C++17 restatement of the reference's proving path, i.e. of what runs
// inside the third-party tools its tests shell out to (snarkjs ^0.7.5 / ffjavascript ^0.2.63,
// un-vendored: /root/reference/package.json:37-46):
//   * witness calculation  -- `node generate_witness.cjs wasm input wtns`
//                             (/root/reference/tests/full_system_simulation.mjs:760-762)
//   * `snarkjs groth16 prove zkey wtns proof public` (:773-775): buildABC1, 3 ifft, odd-coset
//     shift, 3 fft, joinABC, multiExpAffine x5, blinding with r,s (injected, snarkjs uses a CSPRNG)
// It is validated against the pure-Python oracle (oracle/groth16_ref.py, oracle/bn254_ref.py),
// which in turn is pinned by /root/reference/data/test_input_v5.json and the pairing check.
// Arithmetic here is deliberately different from the CUDA product: 4x64-bit limbs with
// unsigned __int128 and Jacobian coordinates (the product uses 8x32-bit limbs and XYZZ).
//
// Parity status: proof bytes are "parity unpinned" against snarkjs itself (no Node.js on this
// image, no zkey/wtns/proof fixture in the reference tree).
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;

// ------------------------------------------------------------------------------------ fields
struct FqP {
  static constexpr u64 M[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  static constexpr u64 INV = 0x87d20782e4866389ULL;  // -q^-1 mod 2^64
  static constexpr u64 R2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
};
struct FrP {
  static constexpr u64 M[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  static constexpr u64 INV = 0xc2e1f593efffffffULL;
  static constexpr u64 R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
};

template <class P>
struct Fp {
  u64 v[4];
  static Fp zero() { return Fp{{0, 0, 0, 0}}; }
  static Fp raw(const u64* p) { Fp r; memcpy(r.v, p, 32); return r; }
  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  bool operator==(const Fp& o) const { return memcmp(v, o.v, 32) == 0; }
  static bool geq_mod(const u64* a) {
    for (int i = 3; i >= 0; i--) {
      if (a[i] > P::M[i]) return true;
      if (a[i] < P::M[i]) return false;
    }
    return true;
  }
  static void sub_mod(u64* a) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)a[i] - P::M[i] - (u64)b;
      a[i] = (u64)d;
      b = (d >> 64) & 1;
    }
  }
  Fp operator+(const Fp& o) const {
    Fp r; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)v[i] + o.v[i]; r.v[i] = (u64)c; c >>= 64; }
    if (c || geq_mod(r.v)) sub_mod(r.v);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r; u128 b = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)v[i] - o.v[i] - (u64)b; r.v[i] = (u64)d; b = (d >> 64) & 1; }
    if (b) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.v[i] + P::M[i]; r.v[i] = (u64)c; c >>= 64; } }
    return r;
  }
  Fp neg() const { return zero() - *this; }
  Fp dbl() const { return *this + *this; }
  Fp operator*(const Fp& o) const {  // Montgomery CIOS, R = 2^256
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 4; j++) { c += (u128)v[j] * o.v[i] + t[j]; t[j] = (u64)c; c >>= 64; }
      c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
      u64 m = t[0] * P::INV;
      c = (u128)m * P::M[0] + t[0]; c >>= 64;
      for (int j = 1; j < 4; j++) { c += (u128)m * P::M[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
      c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
    }
    Fp r; memcpy(r.v, t, 32);
    if (t[4] || geq_mod(r.v)) sub_mod(r.v);
    return r;
  }
  Fp sqr() const { return *this * *this; }
  static Fp from_canonical(const u64* p) { return raw(p) * raw(P::R2); }
  static Fp from_u64(u64 x) { u64 t[4] = {x, 0, 0, 0}; return from_canonical(t); }
  static Fp one() { return from_u64(1); }
  void to_canonical(u64* out) const { u64 o[4] = {1, 0, 0, 0}; Fp r = *this * raw(o); memcpy(out, r.v, 32); }
  Fp pow(const u64* e, int nlimbs) const {
    Fp r = one();
    for (int i = nlimbs * 64 - 1; i >= 0; i--) { r = r.sqr(); if ((e[i / 64] >> (i % 64)) & 1) r = r * *this; }
    return r;
  }
  Fp inv() const {  // Fermat
    u64 e[4]; memcpy(e, P::M, 32); e[0] -= 2;
    return pow(e, 4);
  }
};
typedef Fp<FqP> Fq;
typedef Fp<FrP> Fr;

struct Fq2 {
  Fq a, b;
  static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  bool is_zero() const { return a.is_zero() && b.is_zero(); }
  bool operator==(const Fq2& o) const { return a == o.a && b == o.b; }
  Fq2 operator+(const Fq2& o) const { return {a + o.a, b + o.b}; }
  Fq2 operator-(const Fq2& o) const { return {a - o.a, b - o.b}; }
  Fq2 neg() const { return {a.neg(), b.neg()}; }
  Fq2 dbl() const { return {a.dbl(), b.dbl()}; }
  Fq2 operator*(const Fq2& o) const {
    Fq aa = a * o.a, bb = b * o.b;
    Fq s = (a + b) * (o.a + o.b);
    return {aa - bb, s - aa - bb};
  }
  Fq2 sqr() const { Fq t = a * b; return {(a + b) * (a - b), t.dbl()}; }
  Fq2 inv() const { Fq d = (a.sqr() + b.sqr()).inv(); return {a * d, (b * d).neg()}; }
};

// ------------------------------------------------------------------------------------ curve (Jacobian, a = 0)
template <class F> struct Aff { F x, y; bool inf() const { return x.is_zero() && y.is_zero(); } };
template <class F>
struct Jac {
  F X, Y, Z;
  static Jac infinity() { return {F::one(), F::one(), F::zero()}; }
  bool inf() const { return Z.is_zero(); }
  static Jac from_affine(const Aff<F>& p) { if (p.inf()) return infinity(); return {p.x, p.y, F::one()}; }
  Jac dbl() const {  // dbl-2009-l
    if (inf()) return *this;
    F A = X.sqr(), B = Y.sqr(), C = B.sqr();
    F D = ((X + B).sqr() - A - C).dbl();
    F E = A.dbl() + A, Fv = E.sqr();
    F X3 = Fv - D.dbl();
    F Y3 = E * (D - X3) - C.dbl().dbl().dbl();
    F Z3 = (Y * Z).dbl();
    return {X3, Y3, Z3};
  }
  Jac add(const Jac& o) const {  // add-2007-bl
    if (inf()) return o;
    if (o.inf()) return *this;
    F Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
    F U1 = X * Z2Z2, U2 = o.X * Z1Z1;
    F S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
    if (U1 == U2) { if (S1 == S2) return dbl(); return infinity(); }
    F H = U2 - U1, I = H.dbl().sqr(), J = H * I, r = (S2 - S1).dbl(), V = U1 * I;
    F X3 = r.sqr() - J - V.dbl();
    F Y3 = r * (V - X3) - (S1 * J).dbl();
    F Z3 = ((Z + o.Z).sqr() - Z1Z1 - Z2Z2) * H;
    return {X3, Y3, Z3};
  }
  Jac madd(const Aff<F>& o) const {  // madd-2007-bl
    if (o.inf()) return *this;
    if (inf()) return from_affine(o);
    F Z1Z1 = Z.sqr();
    F U2 = o.x * Z1Z1, S2 = o.y * Z * Z1Z1;
    if (X == U2) { if (Y == S2) return dbl(); return infinity(); }
    F H = U2 - X, HH = H.sqr(), I = HH.dbl().dbl(), J = H * I, r = (S2 - Y).dbl(), V = X * I;
    F X3 = r.sqr() - J - V.dbl();
    F Y3 = r * (V - X3) - (Y * J).dbl();
    F Z3 = (Z + H).sqr() - Z1Z1 - HH;
    return {X3, Y3, Z3};
  }
  Jac neg() const { return {X, Y.neg(), Z}; }
  Aff<F> to_affine() const {
    if (inf()) return {F::zero(), F::zero()};
    F zi = Z.inv(), zi2 = zi.sqr();
    return {X * zi2, Y * zi2 * zi};
  }
  Jac mul(const u64* k) const {  // canonical 256-bit scalar
    Jac r = infinity();
    for (int i = 255; i >= 0; i--) { r = r.dbl(); if ((k[i / 64] >> (i % 64)) & 1) r = r.add(*this); }
    return r;
  }
};
typedef Aff<Fq> G1A; typedef Jac<Fq> G1J; typedef Aff<Fq2> G2A; typedef Jac<Fq2> G2J;

// ------------------------------------------------------------------------------------ threading helper
static void parallel_for(int ntasks, int nthreads, const std::function<void(int)>& fn) {
  if (nthreads <= 1 || ntasks <= 1) { for (int i = 0; i < ntasks; i++) fn(i); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  int nt = nthreads < ntasks ? nthreads : ntasks;
  for (int t = 0; t < nt; t++) th.emplace_back([&] { for (;;) { int i = next.fetch_add(1); if (i >= ntasks) break; fn(i); } });
  for (auto& t : th) t.join();
}

// ------------------------------------------------------------------------------------ Pippenger MSM
static inline u32 get_window(const u64* k, int bit, int c) {
  int limb = bit / 64, off = bit % 64;
  u64 v = k[limb] >> off;
  if (off + c > 64 && limb < 3) v |= k[limb + 1] << (64 - off);
  return (u32)(v & ((1ull << c) - 1));
}

template <class F>
static Jac<F> msm(const Aff<F>* bases, const u64* scalars /* n x 4 canonical */, size_t n, int nthreads) {
  typedef Jac<F> J;
  if (n == 0) return J::infinity();
  int c = n < 32 ? 3 : n < 1024 ? 7 : n < (1 << 14) ? 10 : n < (1 << 18) ? 13 : 16;
  int W = (254 + c - 1) / c;
  int chunks = 1;
  if (nthreads > W) chunks = (nthreads + W - 1) / W;
  size_t csz = (n + chunks - 1) / chunks;
  std::vector<J> partial((size_t)W * chunks, J::infinity());
  parallel_for(W * chunks, nthreads, [&](int task) {
    int wdx = task / chunks, ch = task % chunks;
    size_t lo = ch * csz, hi = lo + csz < n ? lo + csz : n;
    std::vector<J> buckets((size_t)1 << c, J::infinity());
    for (size_t i = lo; i < hi; i++) {
      u32 d = get_window(scalars + 4 * i, wdx * c, c);
      if (d) buckets[d] = buckets[d].madd(bases[i]);
    }
    J run = J::infinity(), acc = J::infinity();
    for (size_t b = ((size_t)1 << c) - 1; b >= 1; b--) { run = run.add(buckets[b]); acc = acc.add(run); }
    partial[task] = acc;
  });
  J res = J::infinity();
  for (int wdx = W - 1; wdx >= 0; wdx--) {
    for (int i = 0; i < c; i++) res = res.dbl();
    for (int ch = 0; ch < chunks; ch++) res = res.add(partial[(size_t)wdx * chunks + ch]);
  }
  return res;
}

// ------------------------------------------------------------------------------------ NTT over Fr
static Fr fr_root(int power) {  // ffjavascript: nqr = 5, w[28] = 5^((r-1)/2^28), w[k] = w[k+1]^2
  u64 e[4]; memcpy(e, FrP::M, 32); e[0] -= 1;
  // (r-1) >> 28
  for (int i = 0; i < 4; i++) e[i] = (e[i] >> 28) | (i < 3 ? e[i + 1] << 36 : 0);
  Fr w = Fr::from_u64(5).pow(e, 4);
  for (int i = 28; i > power; i--) w = w.sqr();
  return w;
}

static void ntt(std::vector<Fr>& a, bool inverse) {
  size_t n = a.size();
  int lg = 0; while (((size_t)1 << lg) < n) lg++;
  Fr w = fr_root(lg);
  if (inverse) w = w.inv();
  for (size_t i = 1, j = 0; i < n; i++) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j |= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  std::vector<Fr> tw(n / 2 ? n / 2 : 1);
  tw[0] = Fr::one();
  for (size_t i = 1; i < n / 2; i++) tw[i] = tw[i - 1] * w;
  for (size_t len = 2; len <= n; len <<= 1) {
    size_t step = n / len;
    for (size_t s = 0; s < n; s += len)
      for (size_t k = 0; k < len / 2; k++) {
        Fr u = a[s + k], v = a[s + k + len / 2] * tw[k * step];
        a[s + k] = u + v; a[s + k + len / 2] = u - v;
      }
  }
  if (inverse) { Fr ni = Fr::from_u64(n).inv(); for (auto& x : a) x = x * ni; }
}

// ------------------------------------------------------------------------------------ iden3 binfile
struct Sections { std::map<u32, std::pair<const uint8_t*, u64>> s; };
static bool parse_sections(const uint8_t* d, size_t len, const char* magic, Sections& out) {
  if (len < 12 || memcmp(d, magic, 4)) return false;
  u32 n; memcpy(&n, d + 8, 4);
  size_t p = 12;
  for (u32 i = 0; i < n; i++) {
    if (p + 12 > len) return false;
    u32 id; u64 l; memcpy(&id, d + p, 4); memcpy(&l, d + p + 4, 8); p += 12;
    if (p + l > len) return false;
    if (!out.s.count(id)) out.s[id] = {d + p, l};
    p += l;
  }
  return true;
}

// ------------------------------------------------------------------------------------ Groth16 prove
struct ZkeyView {
  u32 n_vars, n_public, domain;
  G1A alpha1, beta1, delta1; G2A beta2, delta2;
  const uint8_t *coeffs, *pA, *pB1, *pB2, *pC, *pH; u32 n_coeffs;
};
static bool zkey_view(const uint8_t* d, size_t len, ZkeyView& z) {
  Sections S;
  if (!parse_sections(d, len, "zkey", S)) return false;
  for (u32 id : {1u, 2u, 4u, 5u, 6u, 7u, 8u, 9u}) if (!S.s.count(id)) return false;
  const uint8_t* h = S.s[2].first;
  memcpy(&z.n_vars, h + 72, 4); memcpy(&z.n_public, h + 76, 4); memcpy(&z.domain, h + 80, 4);
  size_t p = 84;
  memcpy(&z.alpha1, h + p, 64); p += 64;
  memcpy(&z.beta1, h + p, 64); p += 64;
  memcpy(&z.beta2, h + p, 128); p += 128;
  p += 128;  // gamma2
  memcpy(&z.delta1, h + p, 64); p += 64;
  memcpy(&z.delta2, h + p, 128);
  z.coeffs = S.s[4].first + 4; memcpy(&z.n_coeffs, S.s[4].first, 4);
  z.pA = S.s[5].first; z.pB1 = S.s[6].first; z.pB2 = S.s[7].first; z.pC = S.s[8].first; z.pH = S.s[9].first;
  return true;
}

static void h_scalars(const ZkeyView& z, const u64* wtns, std::vector<u64>& out) {
  size_t n = z.domain;
  std::vector<Fr> A(n, Fr::zero()), B(n, Fr::zero()), C(n);
  for (u32 i = 0; i < z.n_coeffs; i++) {  // buildABC1: coef is stored as coef*R^2, witness raw
    const uint8_t* e = z.coeffs + 44 * (size_t)i;
    u32 mtx, c, s; memcpy(&mtx, e, 4); memcpy(&c, e + 4, 4); memcpy(&s, e + 8, 4);
    u64 cf[4]; memcpy(cf, e + 12, 32);
    Fr v = Fr::raw(cf) * Fr::raw(wtns + 4 * (size_t)s);
    if (mtx == 0) A[c] = A[c] + v; else B[c] = B[c] + v;
  }
  for (size_t i = 0; i < n; i++) C[i] = A[i] * B[i];
  int lg = 0; while (((size_t)1 << lg) < n) lg++;
  Fr inc = (lg == 28) ? Fr::from_u64(25) : fr_root(lg + 1);
  auto odd = [&](std::vector<Fr>& v) {
    ntt(v, true);
    Fr x = Fr::one();
    for (size_t i = 0; i < n; i++) { v[i] = v[i] * x; x = x * inc; }
    ntt(v, false);
  };
  odd(A); odd(B); odd(C);
  out.resize(4 * n);
  for (size_t i = 0; i < n; i++) (A[i] * B[i] - C[i]).to_canonical(&out[4 * i]);  // joinABC
}

static void put_g1(const G1J& p, uint8_t* out) {
  G1A a = p.to_affine();
  if (p.inf()) { memset(out, 0, 64); return; }
  u64 t[4]; a.x.to_canonical(t); memcpy(out, t, 32); a.y.to_canonical(t); memcpy(out + 32, t, 32);
}
static void put_g2(const G2J& p, uint8_t* out) {
  if (p.inf()) { memset(out, 0, 128); return; }
  G2A a = p.to_affine();
  u64 t[4];
  a.x.a.to_canonical(t); memcpy(out, t, 32); a.x.b.to_canonical(t); memcpy(out + 32, t, 32);
  a.y.a.to_canonical(t); memcpy(out + 64, t, 32); a.y.b.to_canonical(t); memcpy(out + 96, t, 32);
}

static int prove_one(const ZkeyView& z, const u64* w, const u64* r, const u64* s, uint8_t* proof_out,
                     uint8_t* public_out, int nthreads) {
  std::vector<u64> P;
  h_scalars(z, w, P);
  u32 m = z.n_vars, l = z.n_public;
  G1J pi_a = msm<Fq>((const G1A*)z.pA, w, m, nthreads);
  G1J pi_b1 = msm<Fq>((const G1A*)z.pB1, w, m, nthreads);
  G2J pi_b = msm<Fq2>((const G2A*)z.pB2, w, m, nthreads);
  G1J pi_c = msm<Fq>((const G1A*)z.pC, w + 4 * (size_t)(l + 1), m - l - 1, nthreads);
  G1J res_h = msm<Fq>((const G1A*)z.pH, P.data(), z.domain, nthreads);
  G1J d1 = G1J::from_affine(z.delta1);
  pi_a = pi_a.madd(z.alpha1).add(d1.mul(r));
  pi_b = pi_b.madd(z.beta2).add(G2J::from_affine(z.delta2).mul(s));
  pi_b1 = pi_b1.madd(z.beta1).add(d1.mul(s));
  pi_c = pi_c.add(res_h).add(pi_a.mul(s)).add(pi_b1.mul(r));
  Fr rs = Fr::from_canonical(r) * Fr::from_canonical(s);
  u64 nrs[4]; rs.neg().to_canonical(nrs);
  pi_c = pi_c.add(d1.mul(nrs));
  put_g1(pi_a, proof_out); put_g2(pi_b, proof_out + 64); put_g1(pi_c, proof_out + 192);
  if (public_out) memcpy(public_out, w + 4, 32 * (size_t)l);
  return 0;
}

// ------------------------------------------------------------------------------------ witness program (.zkwp)
struct PoseidonK { u32 t, rounds, rp; std::vector<Fr> C, M; };
struct Program {
  u32 n_wires, n_public, n_inputs, n_ops;
  std::vector<u32> ops, lc_off, lc_wire, pos_in;
  std::vector<Fr> lc_coef;
  std::map<u32, PoseidonK> pk;
};
static bool parse_program(const uint8_t* d, size_t len, Program& p) {
  Sections S;
  if (!parse_sections(d, len, "zkwp", S)) return false;
  u32 h[8]; memcpy(h, S.s[1].first, 32);
  p.n_wires = h[0]; p.n_public = h[1]; p.n_inputs = h[2]; p.n_ops = h[3];
  u32 n_lcs = h[4], n_terms = h[5], n_pos = h[6], n_widths = h[7];
  p.ops.resize(5 * (size_t)p.n_ops); memcpy(p.ops.data(), S.s[2].first, 20 * (size_t)p.n_ops);
  p.lc_off.resize(n_lcs + 1); memcpy(p.lc_off.data(), S.s[3].first, 4 * (size_t)(n_lcs + 1));
  p.lc_wire.resize(n_terms); memcpy(p.lc_wire.data(), S.s[4].first, 4 * (size_t)n_terms);
  p.lc_coef.resize(n_terms);
  for (u32 i = 0; i < n_terms; i++) { u64 t[4]; memcpy(t, S.s[5].first + 32 * (size_t)i, 32); p.lc_coef[i] = Fr::from_canonical(t); }
  p.pos_in.resize(n_pos); memcpy(p.pos_in.data(), S.s[6].first, 4 * (size_t)n_pos);
  const uint8_t* q = S.s[7].first;
  for (u32 k = 0; k < n_widths; k++) {
    PoseidonK K; u32 hh[4]; memcpy(hh, q, 16); q += 16;
    K.t = hh[0]; K.rounds = hh[1]; K.rp = hh[2];
    K.C.resize((size_t)K.rounds * K.t); K.M.resize((size_t)K.t * K.t);
    for (auto& c : K.C) { u64 t[4]; memcpy(t, q, 32); q += 32; c = Fr::from_canonical(t); }
    for (auto& c : K.M) { u64 t[4]; memcpy(t, q, 32); q += 32; c = Fr::from_canonical(t); }
    p.pk[K.t] = K;
  }
  return true;
}

static void poseidon_perm(const PoseidonK& K, Fr* st, Fr* trace /* 3 per sbox, may be null */) {
  u32 t = K.t; Fr tmp[17];
  for (u32 r = 0; r < K.rounds; r++) {
    for (u32 i = 0; i < t; i++) st[i] = st[i] + K.C[r * t + i];
    bool full = r < 4 || r >= 4 + K.rp;
    for (u32 i = 0; i < (full ? t : 1); i++) {
      Fr x2 = st[i].sqr(), x4 = x2.sqr(), x5 = x4 * st[i];
      if (trace) { trace[0] = x2; trace[1] = x4; trace[2] = x5; trace += 3; }
      st[i] = x5;
    }
    for (u32 i = 0; i < t; i++) { Fr acc = Fr::zero(); for (u32 j = 0; j < t; j++) acc = acc + K.M[i * t + j] * st[j]; tmp[i] = acc; }
    for (u32 i = 0; i < t; i++) st[i] = tmp[i];
  }
  if (trace) trace[0] = st[0];
}

static void witness_one(const Program& p, const u64* inputs /* canonical */, u64* out /* canonical n_wires x 4 */) {
  std::vector<Fr> w(p.n_wires, Fr::zero());
  w[0] = Fr::one();
  for (u32 i = 0; i < p.n_inputs; i++) w[1 + i] = Fr::from_canonical(inputs + 4 * (size_t)i);
  auto lc = [&](u32 k) { Fr a = Fr::zero(); for (u32 i = p.lc_off[k]; i < p.lc_off[k + 1]; i++) a = a + p.lc_coef[i] * w[p.lc_wire[i]]; return a; };
  for (u32 o = 0; o < p.n_ops; o++) {
    const u32* op = &p.ops[5 * (size_t)o];
    u32 dst = op[1];
    switch (op[0]) {
      case 1: w[dst] = lc(op[2]); break;
      case 2: { Fr v = lc(op[2]) * lc(op[3]); if (op[4] != 0xFFFFFFFFu) v = v + lc(op[4]); w[dst] = v; break; }
      case 3: { u64 c[4]; lc(op[2]).to_canonical(c); for (u32 i = 0; i < op[3]; i++) w[dst + i] = ((c[i / 64] >> (i % 64)) & 1) ? Fr::one() : Fr::zero(); break; }
      case 4: { const PoseidonK& K = p.pk.at(op[2]); Fr st[17]; st[0] = Fr::zero(); for (u32 i = 1; i < K.t; i++) st[i] = w[p.pos_in[op[3] + i - 1]]; poseidon_perm(K, st, &w[dst]); break; }
    }
  }
  for (u32 i = 0; i < p.n_wires; i++) w[i].to_canonical(out + 4 * (size_t)i);
}

// ------------------------------------------------------------------------------------ C entry points (ctypes)
extern "C" {

int zo_fr_mul(const u64* a, const u64* b, u64* out) { (Fr::from_canonical(a) * Fr::from_canonical(b)).to_canonical(out); return 0; }
int zo_fq_mul(const u64* a, const u64* b, u64* out) { (Fq::from_canonical(a) * Fq::from_canonical(b)).to_canonical(out); return 0; }

// bases: affine Montgomery LE (zkey layout); scalars: canonical LE; out: affine canonical LE
int zo_g1_msm(const uint8_t* bases, const uint8_t* scalars, u64 n, uint8_t* out, int nthreads) {
  put_g1(msm<Fq>((const G1A*)bases, (const u64*)scalars, n, nthreads), out); return 0;
}
int zo_g2_msm(const uint8_t* bases, const uint8_t* scalars, u64 n, uint8_t* out, int nthreads) {
  put_g2(msm<Fq2>((const G2A*)bases, (const u64*)scalars, n, nthreads), out); return 0;
}
// out[i] = scalars[i] * G (affine Montgomery LE), the way setup writes zkey points
int zo_g1_mul_gen(const uint8_t* scalars, u64 n, uint8_t* out, int nthreads) {
  G1J g = G1J::from_affine(G1A{Fq::from_u64(1), Fq::from_u64(2)});
  parallel_for((int)((n + 255) / 256), nthreads, [&](int blk) {
    for (u64 i = (u64)blk * 256; i < n && i < (u64)(blk + 1) * 256; i++) {
      G1J p = g.mul((const u64*)scalars + 4 * i);
      G1A a = p.to_affine();
      memcpy(out + 64 * i, &a, 64);
    }
  });
  return 0;
}
int zo_g2_mul_gen(const uint8_t* scalars, u64 n, uint8_t* out, int nthreads) {
  static const u64 G2X0[4] = {0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL};
  static const u64 G2X1[4] = {0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL};
  static const u64 G2Y0[4] = {0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL};
  static const u64 G2Y1[4] = {0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL};
  G2A ga{{Fq::from_canonical(G2X0), Fq::from_canonical(G2X1)}, {Fq::from_canonical(G2Y0), Fq::from_canonical(G2Y1)}};
  G2J g = G2J::from_affine(ga);
  parallel_for((int)((n + 63) / 64), nthreads, [&](int blk) {
    for (u64 i = (u64)blk * 64; i < n && i < (u64)(blk + 1) * 64; i++) {
      G2A a = g.mul((const u64*)scalars + 4 * i).to_affine();
      memcpy(out + 128 * i, &a, 128);
    }
  });
  return 0;
}

int zo_h_scalars(const uint8_t* zkey, u64 zkey_len, const uint8_t* wtns, uint8_t* out /* domain x 32 */) {
  ZkeyView z; if (!zkey_view(zkey, zkey_len, z)) return -1;
  std::vector<u64> P; h_scalars(z, (const u64*)wtns, P);
  memcpy(out, P.data(), 32 * (size_t)z.domain); return 0;
}

// wtns: n_vars x 32 canonical LE; r, s: 32 B canonical; proof_out: 256 B (A|B|C affine canonical); public_out: n_public x 32
int zo_groth16_prove(const uint8_t* zkey, u64 zkey_len, const uint8_t* wtns, const uint8_t* r, const uint8_t* s,
                     uint8_t* proof_out, uint8_t* public_out, int nthreads) {
  ZkeyView z; if (!zkey_view(zkey, zkey_len, z)) return -1;
  return prove_one(z, (const u64*)wtns, (const u64*)r, (const u64*)s, proof_out, public_out, nthreads);
}
// batch: one proof per task, proofs spread over the threads (the best case for CPU throughput)
int zo_groth16_prove_batch(const uint8_t* zkey, u64 zkey_len, const uint8_t* wtns, const uint8_t* rs /* B x 64 */,
                           int B, uint8_t* proofs_out /* B x 256 */, uint8_t* publics_out, int nthreads) {
  ZkeyView z; if (!zkey_view(zkey, zkey_len, z)) return -1;
  parallel_for(B, nthreads, [&](int b) {
    prove_one(z, (const u64*)wtns + 4 * (size_t)z.n_vars * b, (const u64*)(rs + 64 * (size_t)b),
              (const u64*)(rs + 64 * (size_t)b + 32), proofs_out + 256 * (size_t)b,
              publics_out ? publics_out + 32 * (size_t)z.n_public * b : nullptr, 1);
  });
  return 0;
}
int zo_zkey_info(const uint8_t* zkey, u64 zkey_len, u32* out3) {
  ZkeyView z; if (!zkey_view(zkey, zkey_len, z)) return -1;
  out3[0] = z.n_vars; out3[1] = z.n_public; out3[2] = z.domain; return 0;
}

// inputs: B x n_inputs x 32 canonical; out: B x n_wires x 32 canonical
int zo_witness_batch(const uint8_t* prog, u64 prog_len, const uint8_t* inputs, int B, uint8_t* out, int nthreads) {
  Program p; if (!parse_program(prog, prog_len, p)) return -1;
  parallel_for(B, nthreads, [&](int b) {
    witness_one(p, (const u64*)inputs + 4 * (size_t)p.n_inputs * b, (u64*)out + 4 * (size_t)p.n_wires * b);
  });
  return 0;
}

// poseidon hash with constants taken from a program blob (so the C++ path is pinned to the Python oracle)
int zo_poseidon(const uint8_t* prog, u64 prog_len, const uint8_t* inputs, int n_in, uint8_t* out) {
  Program p; if (!parse_program(prog, prog_len, p)) return -1;
  auto it = p.pk.find(n_in + 1); if (it == p.pk.end()) return -2;
  Fr st[17]; st[0] = Fr::zero();
  for (int i = 0; i < n_in; i++) st[i + 1] = Fr::from_canonical((const u64*)inputs + 4 * i);
  poseidon_perm(it->second, st, nullptr);
  st[0].to_canonical((u64*)out); return 0;
}

int zo_ntt(uint8_t* data /* n x 32 canonical, in place */, u64 n, int inverse) {
  std::vector<Fr> a(n);
  for (u64 i = 0; i < n; i++) a[i] = Fr::from_canonical((const u64*)data + 4 * i);
  ntt(a, inverse != 0);
  for (u64 i = 0; i < n; i++) a[i].to_canonical((u64*)data + 4 * i);
  return 0;
}

}  // extern "C"
