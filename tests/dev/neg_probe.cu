// dev probe: bisecting k_vfy_prepare (A.x comes back zero on the device)
#include "../../verifiable-federated-training-with-zero-knowledge-proofs-zk-fl-_b200/csrc/kernels.cuh"
#include <cstdio>
namespace zkrt { void note_launch(const char*) {} bool debug_sync() { return false; } void debug_check(const char*, cudaStream_t) {} }
using namespace zk;
using namespace zkp;
template <int V>
__global__ void prep(PairingConsts k, const uint32_t* proofs, G1P* g1s, uint32_t* flags) {
  size_t b = 0;
  uint32_t pw[64];
  ZK_NOUNROLL for (int i = 0; i < 64; i++) pw[i] = proofs[b * 64 + i];
  bool ok = true;
  if (V & 1) { ZK_NOUNROLL for (int i = 0; i < 8; i++) ok = ok && canonical_lt(pw + 8 * i, false); }
  G1P A = g1_from_canonical(pw), C = g1_from_canonical(pw + 48);
  G2P Bp = g2_from_canonical(pw + 16);
  if (V & 2) ok = ok && g1_on_curve(A, k);
  if (V & 4) ok = ok && g1_on_curve(C, k);
  if (V & 8) ok = ok && g2_on_curve(Bp, k);
  A.y = A.y.neg();
  g1s[0] = A;
  g1s[1] = C;
  flags[0] = ok ? 1u : 0u;
}
template <int V> void run(const PairingConsts& k, uint32_t* d, G1P* o, uint32_t* f) {
  prep<V><<<1, 1>>>(k, d, o, f);
  G1P r[2]; cudaError_t e = cudaMemcpy(r, o, sizeof(r), cudaMemcpyDeviceToHost);
  printf("V=%2d %s  A.x %08x..%08x A.y %08x C.x %08x\n", V, cudaGetErrorString(e), r[0].x.v[0], r[0].x.v[7], r[0].y.v[0], r[1].x.v[0]);
}
int main() {
  uint32_t h[64]; for (int i = 0; i < 64; i++) h[i] = (0x01010101u * (i + 1)) & 0x0fffffffu;
  uint32_t *d, *f; G1P* o; cudaMalloc(&d, 256); cudaMalloc(&f, 16); cudaMalloc(&o, 4 * sizeof(G1P)); cudaMemcpy(d, h, 256, cudaMemcpyHostToDevice);
  PairingConsts k = make_consts();
  G1P A = g1_from_canonical(h);
  printf("host A.x %08x..%08x\n", A.x.v[0], A.x.v[7]);
  run<0>(k, d, o, f); run<1>(k, d, o, f); run<2>(k, d, o, f); run<4>(k, d, o, f); run<8>(k, d, o, f); run<15>(k, d, o, f); run<6>(k, d, o, f);
  return 0;
}
