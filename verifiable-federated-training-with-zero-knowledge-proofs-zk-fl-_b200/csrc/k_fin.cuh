// sm_100a kernels of the blinding / proof assembly; only zkfl.cu includes it.
#pragma once
#include "k_msm.cuh"

namespace zk {

// ================================================================================ K8: blinding / finalisation
// phase 1: fixed-base terms. thread (b, k): k=0 r*delta1, 1 s*delta1, 2 -(r*s)*delta1 (G1) ; k=3 s*delta2 (G2)
// rs: canonical [B][2] (host order: r then s). t_g1: [B][3], t_g2: [B]
ZK_GLOBAL void k_fin_fixed(const G1Affine* __restrict__ tab_d1, const G2Affine* __restrict__ tab_d2, const Fr* __restrict__ rs,
                           uint32_t B, G1Xyzz* __restrict__ t_g1, G2Xyzz* __restrict__ t_g2) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)B * 4) return;
  uint32_t b = (uint32_t)(tid / 4), k = (uint32_t)(tid % 4);
  Fr r = rs[2 * (size_t)b], sv = rs[2 * (size_t)b + 1];
  if (k == 0) t_g1[3 * (size_t)b] = fixed_base_mul<Fq>(tab_d1, r.v);
  else if (k == 1) t_g1[3 * (size_t)b + 1] = fixed_base_mul<Fq>(tab_d1, sv.v);
  else if (k == 2) {
    Fr nrs = (r.to_mont() * sv.to_mont()).neg().from_mont();
    t_g1[3 * (size_t)b + 2] = fixed_base_mul<Fq>(tab_d1, nrs.v);
  } else t_g2[b] = fixed_base_mul<Fq2>(tab_d2, sv.v);
}
// phase 2: thread (b, k): k=0: pi_a = A + alpha + r*delta1, then s*pi_a ; k=1: pi_b1 = B1 + beta1 + s*delta1, then r*pi_b1
// msm_g1: [4][B] = A, B1, C, H results. outputs pis[B][2] (pi_a, pi_b1) and var[B][2] (s*pi_a, r*pi_b1)
ZK_GLOBAL void k_fin_var(VkDev vk, const Fr* __restrict__ rs, uint32_t B, const G1Xyzz* __restrict__ msm_g1,
                         const G1Xyzz* __restrict__ t_g1, G1Xyzz* __restrict__ pis, G1Xyzz* __restrict__ var) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)B * 2) return;
  uint32_t b = (uint32_t)(tid / 2), k = (uint32_t)(tid % 2);
  Fr r = rs[2 * (size_t)b], sv = rs[2 * (size_t)b + 1];
  G1Xyzz p = msm_g1[(size_t)k * B + b];
  xyzz_madd(p, k == 0 ? vk.alpha1 : vk.beta1, false);
  xyzz_add(p, t_g1[3 * (size_t)b + k]);
  pis[tid] = p;
  var[tid] = xyzz_scalar_mul(p, k == 0 ? sv.v : r.v);
}
// phase 3: assemble, normalise, write the 256-byte proof (A | B | C, affine canonical LE)
ZK_GLOBAL void k_fin_write(VkDev vk, uint32_t B, const G1Xyzz* __restrict__ msm_g1, const G2Xyzz* __restrict__ msm_g2,
                           const G1Xyzz* __restrict__ t_g1, const G2Xyzz* __restrict__ t_g2, const G1Xyzz* __restrict__ pis,
                           const G1Xyzz* __restrict__ var, Fq* __restrict__ proofs /* [B][8] */) {
  size_t tid = ZK_TID;
  if (tid >= (size_t)B * 3) return;
  uint32_t b = (uint32_t)(tid / 3), k = (uint32_t)(tid % 3);
  Fq* out = proofs + 8 * (size_t)b;
  if (k == 0) {
    G1Affine a = xyzz_to_affine(pis[2 * (size_t)b]);
    out[0] = a.x.from_mont(); out[1] = a.y.from_mont();
  } else if (k == 1) {
    G2Xyzz p = msm_g2[b];
    xyzz_madd(p, vk.beta2, false);
    xyzz_add(p, t_g2[b]);
    G2Affine a = xyzz_to_affine(p);
    out[2] = a.x.a.from_mont(); out[3] = a.x.b.from_mont(); out[4] = a.y.a.from_mont(); out[5] = a.y.b.from_mont();
  } else {
    G1Xyzz p = msm_g1[2 * (size_t)B + b];
    xyzz_add(p, msm_g1[3 * (size_t)B + b]);
    xyzz_add(p, var[2 * (size_t)b]);
    xyzz_add(p, var[2 * (size_t)b + 1]);
    xyzz_add(p, t_g1[3 * (size_t)b + 2]);
    G1Affine a = xyzz_to_affine(p);
    out[6] = a.x.from_mont(); out[7] = a.y.from_mont();
  }
}

}  // namespace zk
