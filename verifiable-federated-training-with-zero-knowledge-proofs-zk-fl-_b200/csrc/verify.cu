// libzkfl.so: Groth16 verification -- single proof on the host, batches on the GPU (same pairing code, pairing.cuh).
#define ZKFL_FQ2_OPS_AS_ONE_CALL 1   // see bn254.cuh: Fq2 products as one call with interleaved Fq products (latency)
#include "host.h"
#include "k_verify.cuh"
#include "verify_host.h"

extern "C" {
int zkfl_groth16_verify(const uint8_t* alpha1, const uint8_t* beta2, const uint8_t* gamma2, const uint8_t* delta2, const uint8_t* ic,
                        const uint8_t* publics, uint32_t n_public, const uint8_t* proof, int* ok) {
  if (!alpha1 || !beta2 || !gamma2 || !delta2 || !ic || (!publics && n_public) || !proof || !ok) return fail(ZKFL_ERR_ARG, "bad argument");
  int r = zkv::groth16_verify(alpha1, beta2, gamma2, delta2, ic, publics, n_public, proof);
  if (r < 0) return fail(ZKFL_ERR_FORMAT, "malformed verification key (coordinate not reduced or point off the curve)");
  *ok = r;
  return 0;
}
// per-key tables (zkfl_ctx::vk_cache): found or built; *out stays valid until four other keys have been used
static int vk_tables(zkfl_ctx* c, const zkp::PairingConsts& k, const uint8_t* alpha1, const uint8_t* beta2, const uint8_t* gamma2,
                     const uint8_t* delta2, const uint8_t* ic, uint32_t l, const std::vector<G1Affine>& ic_m, const zkp::G1P& alpha,
                     const zkp::G2P& beta, const zkp::G2P& gamma, const zkp::G2P& delta, zkfl_ctx::VkCacheEntry** out) {
  std::vector<uint8_t> key(64 + 3 * 128 + 64 * ((size_t)l + 1));
  memcpy(key.data(), alpha1, 64); memcpy(key.data() + 64, beta2, 128); memcpy(key.data() + 192, gamma2, 128);
  memcpy(key.data() + 320, delta2, 128); memcpy(key.data() + 448, ic, 64 * ((size_t)l + 1));
  zkfl_ctx::VkCacheEntry* e = nullptr;
  for (auto& x : c->vk_cache) if (x.key == key) e = &x;
  if (!e) {
    e = &c->vk_cache[0];
    for (auto& x : c->vk_cache) if (x.stamp < e->stamp) e = &x;
    e->key.clear();
    Stage st(c, "verify_key_tables");
    TRY(e->tabs.reserve(((size_t)l + 2) * 8192 * sizeof(G1Affine)));
    TRY(e->lines.reserve((size_t)3 * zkp::kMillerSteps * sizeof(zkp::LineRec)));
    for (uint32_t i = 0; i <= l + 1; i++) {
      const G1Affine base = i <= l ? ic_m[i] : zkp::g1_to_affine(alpha);
      ZK_LAUNCH(k_fixed_base_table<Fq>, 32 * 256, 64, c->stream, base, e->tabs.as<G1Affine>() + (size_t)i * 8192);
    }
    ZK_LAUNCH(k_vfy_lines, 3, 32, c->stream, k, (const zkp::G2P*)nullptr, (const uint32_t*)nullptr, 0u, gamma, delta, beta, e->lines.as<zkp::LineRec>());
    CU(cudaGetLastError());
    e->key = key;
  }
  e->stamp = ++c->vk_stamp;
  *out = e;
  return 0;
}
// per-proof verdicts, lane-cooperative kernels: g1s = [3B + 1] G1 points, line tables of the B proofs + gamma, delta, beta
static int verify_each_coop(zkfl_ctx* c, const zkp::PairingConsts& k, const zkfl_ctx::VkCacheEntry* vkc, const zkp::G2P& beta,
                            const zkp::G2P& gamma, const zkp::G2P& delta, const zkp::G1P& alpha, uint32_t l, uint32_t Bu, bool lines_ready) {
  {
    Stage st(c, "verify_prepare");
    ZK_LAUNCH(k_vfy_ic_mul, (size_t)Bu * l, 64, c->stream, c->v_ic.as<G1Affine>(), c->v_pub.as<Fr>(), l, Bu, c->v_t.as<G1Xyzz>(),
              (const G1Affine*)vkc->tabs.as<G1Affine>());
    ZK_LAUNCH(k_vfy_prepare, Bu, 32, c->stream, k, c->v_ic.as<G1Affine>(), c->v_pub.as<Fr>(), l, Bu, c->v_proofs.as<uint32_t>(),
              c->v_t.as<G1Xyzz>(), c->v_g1.as<zkp::G1P>(), c->v_g2.as<zkp::G2P>(), c->v_flags.as<uint32_t>());
    ZK_LAUNCH(k_vfy_put_g1, 1, 32, c->stream, alpha, c->v_g1.as<zkp::G1P>() + (size_t)3 * Bu);
    if (!lines_ready) {   // the proofs' tables; the key's three come from the cache (k_vfy_lines would redo them: t >= B)
      ZK_LAUNCH(k_vfy_lines, (size_t)Bu, 32, c->stream, k, c->v_g2.as<zkp::G2P>(), c->v_flags.as<uint32_t>(), Bu, gamma, delta, beta,
                c->v_lines.as<zkp::LineRec>());
      CU(cudaMemcpyAsync(c->v_lines.as<zkp::LineRec>() + (size_t)Bu * zkp::kMillerSteps, vkc->lines.p,
                         (size_t)3 * zkp::kMillerSteps * sizeof(zkp::LineRec), cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  {
    Stage st(c, "verify_miller");
    ZK_LAUNCH(k_vfy_miller_coop, ((size_t)3 * Bu + 1) * 8, 64, c->stream, c->v_g1.as<zkp::G1P>(), c->v_lines.as<zkp::LineRec>(),
              c->v_flags.as<uint32_t>(), Bu, 3 * Bu + 1, 0, c->v_f.as<zkp::T12>());
  }
  {
    Stage st(c, "verify_final_exp");
    ZK_LAUNCH(k_vfy_final_coop, (size_t)Bu * 8, 64, c->stream, k, c->v_f.as<zkp::T12>(), 3u, 3u, 3 * Bu, c->v_flags.as<uint32_t>(), Bu,
              c->v_ok.as<int32_t>());
  }
  CU(cudaGetLastError());
  return 0;
}
// one combined check under random weights; *passed = 1 when the batch verifies and every B lies in the r-torsion subgroup
// (then v_ok holds the verdicts), 0 when the caller has to fall back to the per-proof form (the line tables stay valid)
static int verify_rlc(zkfl_ctx* c, const zkp::PairingConsts& k, const zkfl_ctx::VkCacheEntry* vkc, uint32_t l, uint32_t Bu, int* passed) {
  // 128 random bits per proof
  std::vector<uint32_t> rho((size_t)Bu * 8, 0u);
  {
    FILE* f = fopen("/dev/urandom", "rb");
    if (!f) return fail(ZKFL_ERR_ARG, "cannot open /dev/urandom");
    std::vector<uint32_t> raw((size_t)Bu * 4);
    const bool ok = fread(raw.data(), 16, Bu, f) == Bu;
    fclose(f);
    if (!ok) return fail(ZKFL_ERR_ARG, "urandom read failed");
    for (uint32_t b = 0; b < Bu; b++) {
      for (int i = 0; i < 4; i++) rho[(size_t)b * 8 + i] = raw[(size_t)b * 4 + i];
      if (!(rho[(size_t)b * 8] | rho[(size_t)b * 8 + 1] | rho[(size_t)b * 8 + 2] | rho[(size_t)b * 8 + 3])) rho[(size_t)b * 8] = 1;   // never zero
    }
  }
  const uint32_t nch = (Bu + 63) / 64;
  TRY(c->v_rho.reserve((size_t)Bu * 32));
  TRY(c->v_cps.reserve((size_t)Bu * sizeof(G1Xyzz)));
  TRY(c->v_sum[0].reserve(((size_t)Bu / 16 + 2) * sizeof(G1Xyzz)));
  TRY(c->v_sum[1].reserve(((size_t)Bu / 256 + 2) * sizeof(G1Xyzz)));
  TRY(c->v_s.reserve(((size_t)l + 1) * 32));
  TRY(c->v_spart.reserve(((size_t)l + 1) * nch * 32));
  TRY(c->v_tmul.reserve(((size_t)l + 2) * sizeof(G1Xyzz)));
  TRY(c->v_sub.reserve((size_t)Bu * 4));
  TRY(c->v_tree[0].reserve(((size_t)Bu / 8 + 2) * sizeof(zkp::T12)));
  TRY(c->v_tree[1].reserve(((size_t)Bu / 64 + 2) * sizeof(zkp::T12)));
  TRY(c->v_misc.reserve(64));
  CU(cudaMemcpyAsync(c->v_rho.p, rho.data(), (size_t)Bu * 32, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(c->v_misc.p, 0, 64, c->stream));
  {
    Stage st(c, "verify_prepare");
    ZK_LAUNCH(k_vfy_check, Bu, 64, c->stream, k, c->v_pub.as<Fr>(), l, Bu, c->v_proofs.as<uint32_t>(), c->v_g2.as<zkp::G2P>(),
              c->v_flags.as<uint32_t>());
    ZK_LAUNCH(k_vfy_rlc_scalars_part, ((size_t)l + 1) * nch, 64, c->stream, c->v_rho.as<Fr>(), c->v_pub.as<Fr>(), c->v_flags.as<uint32_t>(), l, Bu,
              c->v_spart.as<Fr>());
    ZK_LAUNCH(k_vfy_rlc_scalars_sum, (size_t)l + 1, 32, c->stream, c->v_spart.as<Fr>(), l, nch, c->v_s.as<Fr>());
    const size_t wB = ((size_t)Bu + 31) / 32, wF = ((size_t)l + 2 + 31) / 32;
    ZK_LAUNCH(k_vfy_rlc_stage1, (4 * wB + wF) * 32, 32, c->stream, k, c->v_proofs.as<uint32_t>(), c->v_rho.as<Fr>(), c->v_flags.as<uint32_t>(), Bu,
              l, c->v_g2.as<zkp::G2P>(), (const G1Affine*)vkc->tabs.as<G1Affine>(), c->v_s.as<Fr>(), c->v_lines.as<zkp::LineRec>(),
              c->v_g1.as<zkp::G1P>(), c->v_cps.as<G1Xyzz>(), c->v_sub.as<uint32_t>(), c->v_tmul.as<G1Xyzz>());
    CU(cudaMemcpyAsync(c->v_lines.as<zkp::LineRec>() + (size_t)Bu * zkp::kMillerSteps, vkc->lines.p,
                       (size_t)3 * zkp::kMillerSteps * sizeof(zkp::LineRec), cudaMemcpyDeviceToDevice, c->stream));
    // sum of the rho_b C_b: fan-in-16 tree
    const G1Xyzz* in = c->v_cps.as<G1Xyzz>();
    uint32_t n = Bu; int pp = 0;
    while (n > 1) {
      ZK_LAUNCH(k_vfy_sum_g1, ((size_t)n + 15) / 16, 32, c->stream, in, n, 16u, c->v_sum[pp].as<G1Xyzz>());
      in = c->v_sum[pp].as<G1Xyzz>(); n = (n + 15) / 16; pp ^= 1;
    }
    ZK_LAUNCH(k_vfy_rlc_points, 3, 32, c->stream, c->v_tmul.as<G1Xyzz>(), l, in, Bu, c->v_g1.as<zkp::G1P>());
  }
  const zkp::T12* fin = c->v_f.as<zkp::T12>();
  uint32_t nf = Bu + 3;
  {
    Stage st(c, "verify_miller");
    ZK_LAUNCH(k_vfy_miller_coop, ((size_t)Bu + 3) * 8, 64, c->stream, c->v_g1.as<zkp::G1P>(), c->v_lines.as<zkp::LineRec>(),
              c->v_flags.as<uint32_t>(), Bu, Bu + 3, 1, c->v_f.as<zkp::T12>());
    int pp = 0;
    while (nf > 8) {     // product tree, fan-in 8
      ZK_LAUNCH(k_vfy_prod_coop, (((size_t)nf + 7) / 8) * 8, 64, c->stream, fin, nf, 8u, c->v_tree[pp].as<zkp::T12>());
      fin = c->v_tree[pp].as<zkp::T12>(); nf = (nf + 7) / 8; pp ^= 1;
    }
  }
  {
    Stage st(c, "verify_final_exp");
    ZK_LAUNCH(k_vfy_final_coop, 8, 32, c->stream, k, fin, nf, nf, 0xFFFFFFFFu, (const uint32_t*)nullptr, 1u, c->v_misc.as<int32_t>() + 1);
    ZK_LAUNCH(k_vfy_rlc_verdicts, Bu, 64, c->stream, c->v_flags.as<uint32_t>(), c->v_sub.as<uint32_t>(), Bu, c->v_ok.as<int32_t>(),
              c->v_misc.as<uint32_t>());
  }
  CU(cudaGetLastError());
  uint32_t h[2] = {0, 0};     // h[0]: some B outside the subgroup; h[1]: the combined check holds
  CU(cudaMemcpyAsync(h, c->v_misc.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  *passed = (h[1] == 1 && h[0] == 0) ? 1 : 0;
  return 0;
}
int zkfl_groth16_verify_batch(zkfl_ctx* c, const uint8_t* alpha1, const uint8_t* beta2, const uint8_t* gamma2, const uint8_t* delta2,
                              const uint8_t* ic, uint32_t n_public, const uint8_t* publics, const uint8_t* proofs, int B, int32_t* ok) {
  if (!c || !alpha1 || !beta2 || !gamma2 || !delta2 || !ic || (!publics && n_public) || !proofs || !ok || B < 0)
    return fail(ZKFL_ERR_ARG, "bad argument");
  if (B == 0) return 0;
  CU(cudaSetDevice(c->device));
  const zkp::PairingConsts& k = zkv::consts();
  // verification key: validated by the helper the host verifier uses, decoded on the host (a handful of points)
  if (!zkv::vkey_well_formed(alpha1, beta2, gamma2, delta2, ic, n_public)) return fail(ZKFL_ERR_FORMAT, "malformed verification key (coordinate not reduced or point off the curve)");
  uint32_t w[32];
  memcpy(w, alpha1, 64); const zkp::G1P alpha = zkp::g1_from_canonical(w);
  memcpy(w, beta2, 128); const zkp::G2P beta = zkp::g2_from_canonical(w);
  memcpy(w, gamma2, 128); const zkp::G2P gamma = zkp::g2_from_canonical(w);
  memcpy(w, delta2, 128); const zkp::G2P delta = zkp::g2_from_canonical(w);
  std::vector<G1Affine> ic_m(n_public + 1);
  for (uint32_t i = 0; i <= n_public; i++) {
    memcpy(w, ic + 64 * (size_t)i, 64);
    ic_m[i] = zkp::g1_to_affine(zkp::g1_from_canonical(w));
  }
  const uint32_t l = n_public, Bu = (uint32_t)B;
  // forms: ZKFL_VERIFY_FLAT=1 -- the first (thread-per-proof, flat basis) kernels; ZKFL_VERIFY_COOP=0 -- thread-per-proof in the tower
  // view; default -- lane-cooperative kernels, with the random-linear-combination check in front for batches (ZKFL_VERIFY_RLC=0: off)
  const bool flat = env_u32("ZKFL_VERIFY_FLAT", 0) != 0;
  const bool coop = !flat && env_u32("ZKFL_VERIFY_COOP", 1) != 0;
  const bool rlc = coop && Bu >= env_u32("ZKFL_VERIFY_RLC_MIN", 4) && env_u32("ZKFL_VERIFY_RLC", 1) != 0;
  TRY(c->v_ic.reserve(ic_m.size() * sizeof(G1Affine)));
  TRY(c->v_pub.reserve((size_t)Bu * (l ? l : 1) * sizeof(Fr)));
  TRY(c->v_proofs.reserve((size_t)Bu * 256));
  TRY(c->v_t.reserve((size_t)Bu * (l ? l : 1) * sizeof(G1Xyzz)));
  TRY(c->v_g1.reserve(((size_t)Bu * 3 + 4) * sizeof(zkp::G1P)));
  TRY(c->v_g2.reserve((size_t)Bu * sizeof(zkp::G2P)));
  TRY(c->v_flags.reserve((size_t)Bu * 4));
  TRY(c->v_f.reserve(((size_t)3 * Bu + 4) * sizeof(zkp::F12)));
  TRY(c->v_halves.reserve((size_t)2 * Bu * sizeof(zkp::F12)));
  TRY(c->v_ok.reserve((size_t)Bu * 4));
  if (coop) TRY(c->v_lines.reserve(((size_t)Bu + 3) * zkp::kMillerSteps * sizeof(zkp::LineRec)));
  CU(cudaMemcpyAsync(c->v_ic.p, ic_m.data(), ic_m.size() * sizeof(G1Affine), cudaMemcpyHostToDevice, c->stream));
  if (l) CU(cudaMemcpyAsync(c->v_pub.p, publics, (size_t)Bu * l * 32, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->v_proofs.p, proofs, (size_t)Bu * 256, cudaMemcpyHostToDevice, c->stream));
  if (coop) {
    int passed = 0;
    zkfl_ctx::VkCacheEntry* vkc = nullptr;
    TRY(vk_tables(c, k, alpha1, beta2, gamma2, delta2, ic, l, ic_m, alpha, beta, gamma, delta, &vkc));
    if (rlc) TRY(verify_rlc(c, k, vkc, l, Bu, &passed));
    if (!passed) TRY(verify_each_coop(c, k, vkc, beta, gamma, delta, alpha, l, Bu, rlc));
    c->v_last_rlc = passed;
    CU(cudaMemcpyAsync(ok, c->v_ok.p, (size_t)Bu * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
  }
  {
    Stage st(c, "verify_prepare");
    ZK_LAUNCH(k_vfy_ic_mul, (size_t)Bu * l, 64, c->stream, c->v_ic.as<G1Affine>(), c->v_pub.as<Fr>(), l, Bu, c->v_t.as<G1Xyzz>(),
              (const G1Affine*)nullptr);
    ZK_LAUNCH(k_vfy_prepare, Bu, 32, c->stream, k, c->v_ic.as<G1Affine>(), c->v_pub.as<Fr>(), l, Bu, c->v_proofs.as<uint32_t>(),
              c->v_t.as<G1Xyzz>(), c->v_g1.as<zkp::G1P>(), c->v_g2.as<zkp::G2P>(), c->v_flags.as<uint32_t>());
  }
  {
    Stage st(c, "verify_miller");
    ZK_LAUNCH(k_vfy_miller, (size_t)3 * Bu + 1, 32, c->stream, k, beta, gamma, delta, alpha, c->v_g1.as<zkp::G1P>(),
              c->v_g2.as<zkp::G2P>(), Bu, c->v_f.as<zkp::F12>(), c->v_flags.as<uint32_t>(), (int)flat);
  }
  if (!flat) {
    Stage st(c, "verify_final_exp");
    ZK_LAUNCH(k_vfy_final_tower, Bu, 32, c->stream, k, c->v_f.as<zkp::F12>(), c->v_flags.as<uint32_t>(), Bu, c->v_ok.as<int32_t>());
  } else {   // cross-check knob: the inversion-free two-power form in the flat basis
    Stage st(c, "verify_final_exp");
    ZK_LAUNCH(k_vfy_final, ((size_t)Bu + 31) / 32 * 64, 64, c->stream, k, c->v_f.as<zkp::F12>(), c->v_flags.as<uint32_t>(), Bu,
              c->v_halves.as<zkp::F12>());
    ZK_LAUNCH(k_vfy_compare, Bu, 64, c->stream, c->v_halves.as<zkp::F12>(), c->v_flags.as<uint32_t>(), Bu, c->v_ok.as<int32_t>());
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ok, c->v_ok.p, (size_t)Bu * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
int zkfl_debug_pairing_selftest(void) { return zkv::pairing_selftest(); }
// dev / test hook: copies a named workspace buffer of the batch verifier to the host (intermediate values of the last call)
int zkfl_debug_read(zkfl_ctx* c, const char* name, void* out, size_t bytes) {
  if (!c || !name || !out) return fail(ZKFL_ERR_ARG, "bad argument");
  const std::string n(name);
  if (n == "v_last_rlc" && bytes == 4) { memcpy(out, &c->v_last_rlc, 4); return 0; }
  const DevBuf* b = n == "v_f" ? &c->v_f : n == "v_halves" ? &c->v_halves : n == "v_flags" ? &c->v_flags : n == "v_g1" ? &c->v_g1
                  : n == "v_g2" ? &c->v_g2 : n == "v_t" ? &c->v_t : n == "v_misc" ? &c->v_misc : n == "v_sub" ? &c->v_sub : n == "v_s" ? &c->v_s
                  : n == "v_rho" ? &c->v_rho : n == "v_tmul" ? &c->v_tmul : n == "v_cps" ? &c->v_cps : nullptr;
  if (!b || bytes > b->cap) return fail(ZKFL_ERR_ARG, "unknown buffer or size");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(out, b->p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
}  // extern "C"
