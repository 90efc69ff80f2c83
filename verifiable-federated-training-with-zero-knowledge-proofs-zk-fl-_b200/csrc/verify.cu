// libzkfl.so: Groth16 verification -- single proof on the host, batches on the GPU (same pairing code, pairing.cuh).
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
  TRY(c->v_ic.reserve(ic_m.size() * sizeof(G1Affine)));
  TRY(c->v_pub.reserve((size_t)Bu * (l ? l : 1) * sizeof(Fr)));
  TRY(c->v_proofs.reserve((size_t)Bu * 256));
  TRY(c->v_t.reserve((size_t)Bu * (l ? l : 1) * sizeof(G1Xyzz)));
  TRY(c->v_g1.reserve((size_t)Bu * 3 * sizeof(zkp::G1P)));
  TRY(c->v_g2.reserve((size_t)Bu * sizeof(zkp::G2P)));
  TRY(c->v_flags.reserve((size_t)Bu * 4));
  TRY(c->v_f.reserve(((size_t)3 * Bu + 1) * sizeof(zkp::F12)));
  TRY(c->v_halves.reserve((size_t)2 * Bu * sizeof(zkp::F12)));
  TRY(c->v_ok.reserve((size_t)Bu * 4));
  CU(cudaMemcpyAsync(c->v_ic.p, ic_m.data(), ic_m.size() * sizeof(G1Affine), cudaMemcpyHostToDevice, c->stream));
  if (l) CU(cudaMemcpyAsync(c->v_pub.p, publics, (size_t)Bu * l * 32, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->v_proofs.p, proofs, (size_t)Bu * 256, cudaMemcpyHostToDevice, c->stream));
  {
    Stage st(c, "verify_prepare");
    ZK_LAUNCH(k_vfy_ic_mul, (size_t)Bu * l, 64, c->stream, c->v_ic.as<G1Affine>(), c->v_pub.as<Fr>(), l, Bu, c->v_t.as<G1Xyzz>());
    ZK_LAUNCH(k_vfy_prepare, Bu, 32, c->stream, k, c->v_ic.as<G1Affine>(), c->v_pub.as<Fr>(), l, Bu, c->v_proofs.as<uint32_t>(),
              c->v_t.as<G1Xyzz>(), c->v_g1.as<zkp::G1P>(), c->v_g2.as<zkp::G2P>(), c->v_flags.as<uint32_t>());
  }
  {
    Stage st(c, "verify_miller");
    ZK_LAUNCH(k_vfy_miller, (size_t)3 * Bu + 1, 32, c->stream, k, beta, gamma, delta, alpha, c->v_g1.as<zkp::G1P>(),
              c->v_g2.as<zkp::G2P>(), Bu, c->v_f.as<zkp::F12>(), c->v_flags.as<uint32_t>(), (int)env_u32("ZKFL_VERIFY_FLAT", 0));
  }
  if (!env_u32("ZKFL_VERIFY_FLAT", 0)) {
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
  const DevBuf* b = n == "v_f" ? &c->v_f : n == "v_halves" ? &c->v_halves : n == "v_flags" ? &c->v_flags : n == "v_g1" ? &c->v_g1
                  : n == "v_g2" ? &c->v_g2 : n == "v_t" ? &c->v_t : nullptr;
  if (!b || bytes > b->cap) return fail(ZKFL_ERR_ARG, "unknown buffer or size");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(out, b->p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
}  // extern "C"
