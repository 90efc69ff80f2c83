// libzkfl.so: batched witness evaluation (W1), constraint check, sparse A.w / B.w (K1) and the H polynomial (K2-K5).
#include "host.h"
#include "k_witness.cuh"

int zk_aos_to_soa(zkfl_ctx* c, const Fr* src, Fr* dst, uint32_t n_elem, uint32_t B, uint32_t dst_elem_off) {
  ZK_LAUNCH(k_aos_to_soa, (size_t)n_elem * B, 256, c->stream, src, dst, n_elem, B, dst_elem_off);
  CU(cudaGetLastError());
  return 0;
}
int zk_soa_to_aos(zkfl_ctx* c, const Fr* src, Fr* dst, uint32_t n_elem, uint32_t B) {
  ZK_LAUNCH(k_soa_to_aos, (size_t)n_elem * B, 256, c->stream, src, dst, n_elem, B);
  CU(cudaGetLastError());
  return 0;
}
int zk_gather_wires(zkfl_ctx* c, const Fr* w, const uint32_t* wires, uint32_t n_sel, uint32_t B, Fr* out) {
  ZK_LAUNCH(k_gather_wires, (size_t)n_sel * B, 256, c->stream, w, wires, n_sel, B, out);
  CU(cudaGetLastError());
  return 0;
}

int run_witness(zkfl_ctx* c, const zkfl_circuit* circ, const uint8_t* inputs_host, uint32_t B) {
  Stage st(c, "witness");
  TRY(c->w.reserve((size_t)circ->n_wires * B * sizeof(Fr)));
  if (inputs_host) {
    TRY(c->stage_in.reserve((size_t)circ->n_inputs * B * sizeof(Fr)));
    CU(cudaMemcpyAsync(c->stage_in.p, inputs_host, (size_t)circ->n_inputs * B * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  }
  TRY(zk_aos_to_soa(c, c->stage_in.as<Fr>(), c->w.as<Fr>(), circ->n_inputs, B, 1u));
  ZK_LAUNCH(k_witness_init, B, 128, c->stream, c->w.as<Fr>(), B);
  for (size_t k = 0; k + 1 < circ->level_off.size(); k++) {
    uint32_t lo = circ->level_off[k], hi = circ->level_off[k + 1];
#ifndef ZKFL_EMUL
    // few instances: one warp per (op, instance), Poseidon state across the lanes (ZKFL_WITNESS_COOP = 0 never, 1 always)
    const uint32_t coop = env_u32("ZKFL_WITNESS_COOP", 2);
    if (coop == 1 || (coop == 2 && B <= 32)) {
      ZK_LAUNCH(k_witness_level_coop, (size_t)(hi - lo) * B * 32, 128, c->stream, circ->dev, c->w.as<Fr>(), B, lo, hi);
      continue;
    }
#endif
    ZK_LAUNCH(k_witness_level, (size_t)(hi - lo) * B, 64, c->stream, circ->dev, c->w.as<Fr>(), B, lo, hi);
  }
  CU(cudaGetLastError());
  c->w_wires = circ->n_wires; c->w_B = B;
  return 0;
}

int check_r1cs_device(zkfl_ctx* c, const zkfl_r1cs* r, uint32_t B, uint32_t* first_bad) {
  Stage st(c, "r1cs_check");
  TRY(c->bad.reserve((size_t)B * 4));
  CU(cudaMemsetAsync(c->bad.p, 0xFF, (size_t)B * 4, c->stream));
  ZK_LAUNCH(k_r1cs_check, (size_t)r->n_constraints * B, 128, c->stream, r->A.dev(), r->B.dev(), r->C.dev(), c->w.as<Fr>(),
            r->n_constraints, B, c->bad.as<uint32_t>());
  std::vector<uint32_t> host(B);
  CU(cudaMemcpyAsync(host.data(), c->bad.p, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  int bad = 0;
  for (uint32_t b = 0; b < B; b++) { if (first_bad) first_bad[b] = host[b]; if (host[b] != 0xFFFFFFFFu) bad++; }
  if (bad) return fail(ZKFL_ERR_ASSERT, "Assert Failed: " + std::to_string(bad) + " of " + std::to_string(B) + " witnesses violate a constraint");
  return 0;
}

static int chk_reserve(zkfl_ctx* c, uint32_t B) {
  const size_t need = (size_t)B + 1;
  if (need <= c->chk_cap) return 0;
  if (c->chk_host) cudaFreeHost(c->chk_host);
  c->chk_host = nullptr; c->chk_cap = 0;
  void* p = nullptr;
  if (cudaMallocHost(&p, need * 4) != cudaSuccess) return fail(ZKFL_ERR_NOMEM, "cudaMallocHost failed");
  c->chk_host = (uint32_t*)p; c->chk_cap = need;
  c->chk_host[0] = 0;
  return 0;
}
// reserve + reset the per-instance verdicts (before the checking kernel), then queue their copy to pinned memory (after it)
static int check_prepare(zkfl_ctx* c, uint32_t B) {
  const bool keep_flags = c->chk_wtns;
  const uint32_t flags = keep_flags ? c->chk_host[0] : 0;
  TRY(chk_reserve(c, B));
  if (keep_flags) c->chk_host[0] = flags;
  TRY(c->bad.reserve(((size_t)B + 1) * 4));
  CU(cudaMemsetAsync(c->bad.as<uint32_t>() + 1, 0xFF, (size_t)B * 4, c->stream));
  return 0;
}
static int check_collect(zkfl_ctx* c, uint32_t B) {
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->chk_host + 1, c->bad.as<uint32_t>() + 1, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
  c->chk_B = B;
  return 0;
}
int check_r1cs_launch(zkfl_ctx* c, const zkfl_r1cs* r, uint32_t B) {
  Stage st(c, "r1cs_check");
  TRY(check_prepare(c, B));
  ZK_LAUNCH(k_r1cs_check, (size_t)r->n_constraints * B, 128, c->stream, r->A.dev(), r->B.dev(), r->C.dev(), c->w.as<Fr>(),
            r->n_constraints, B, c->bad.as<uint32_t>() + 1);
  return check_collect(c, B);
}
int check_wtns_launch(zkfl_ctx* c, uint32_t n_wires, uint32_t B) {
  TRY(chk_reserve(c, c->chk_B > B ? c->chk_B : B));
  TRY(c->bad.reserve(((size_t)(c->chk_B > B ? c->chk_B : B) + 1) * 4));
  CU(cudaMemsetAsync(c->bad.p, 0, 4, c->stream));
  ZK_LAUNCH(k_wtns_validate, (size_t)n_wires * B, 256, c->stream, c->w.as<Fr>(), n_wires, B, c->bad.as<uint32_t>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->chk_host, c->bad.p, 4, cudaMemcpyDeviceToHost, c->stream));
  c->chk_wtns = true;
  return 0;
}
// call after the stream has been synchronised
int checks_result(zkfl_ctx* c, uint32_t* first_bad, uint32_t cap) {
  int rc = 0;
  if (c->chk_wtns) {
    c->chk_wtns = false;
    const uint32_t f = c->chk_host[0];
    if (f & 1u) rc = fail(ZKFL_ERR_ARG, "witness element not reduced mod r");
    else if (f & 2u) rc = fail(ZKFL_ERR_ARG, "witness wire 0 is not 1");
  }
  if (c->chk_B) {
    const uint32_t B = c->chk_B;
    c->chk_B = 0;
    uint32_t bad = 0;
    for (uint32_t b = 0; b < B; b++) { if (first_bad && b < cap) first_bad[b] = c->chk_host[1 + b]; if (c->chk_host[1 + b] != 0xFFFFFFFFu) bad++; }
    if (bad && !rc) rc = fail(ZKFL_ERR_ASSERT, "Assert Failed: " + std::to_string(bad) + " of " + std::to_string(B) + " witnesses violate a constraint");
  }
  return rc;
}

// A.w, B.w, C = A o B over the constraint domain, then 3 iNTT -> odd-coset shift -> 3 NTT -> A*B - C (snarkjs: buildABC1,
// ifft x3, batchApplyKey, fft x3, joinABC); the witness is in c->w, the H-MSM scalars (canonical) land in c->hsc
int run_h_poly(zkfl_ctx* c, const zkfl_zkey* z, uint32_t B, const zkfl_r1cs* check) {
  const uint32_t n = z->domain;
  Fr* w = c->w.as<Fr>();
  TRY(c->abc.reserve(3 * (size_t)n * B * sizeof(Fr)));
  TRY(c->hsc.reserve((size_t)n * B * sizeof(Fr)));
  Fr* abc = c->abc.as<Fr>();
  if (check) {       // full-prove: the constraint check rides on the A.w / B.w products (deferred verdict, see checks_result)
    if (check->n_wires != z->n_vars || check->n_constraints > n) return fail(ZKFL_ERR_ARG, "r1cs does not match the proving key");
    Stage st(c, "build_abc");
    TRY(check_prepare(c, B));
    ZK_LAUNCH(k_build_abc_check, (size_t)n * B, 128, c->stream, z->A.dev(), z->B.dev(), check->C.dev(), w, abc, n, B, check->n_constraints,
              c->bad.as<uint32_t>() + 1);
    TRY(check_collect(c, B));
  } else {
    Stage st(c, "build_abc");
    ZK_LAUNCH(k_build_abc, (size_t)n * B, 128, c->stream, z->A.dev(), z->B.dev(), w, abc, n, B);
  }
  {
    Stage st(c, "ntt");
    // inverse transform (DIF, natural -> bit-reversed), 3 stages per pass; the last pass also applies n^-1 * w_2n^bitrev(p)
    auto radix = [&](int K, uint32_t half, int dif, const Fr* twd, const Fr* scale) {
      size_t threads = (size_t)3 * (n >> K) * B;
      if (K == 3) ZK_LAUNCH(k_ntt_radix<3>, threads, 128, c->stream, abc, twd, scale, n, B, 3u, half, dif);
      else if (K == 2) ZK_LAUNCH(k_ntt_radix<2>, threads, 128, c->stream, abc, twd, scale, n, B, 3u, half, dif);
      else ZK_LAUNCH(k_ntt_radix<1>, threads, 256, c->stream, abc, twd, scale, n, B, 3u, half, dif);
    };
    const int lg = (int)z->log_n;
    for (int done = 0; done < lg;) {   // DIF: stage t has half = n >> (t + 1)
      int K = lg - done >= 3 ? 3 : lg - done;
      bool last = done + K == lg;
      radix(K, n >> (done + 1), 1, z->tw_inv.as<Fr>(), last ? z->coset.as<Fr>() : nullptr);
      done += K;
    }
    for (int done = 0; done < lg;) {   // DIT: stage t has half = 1 << t
      int K = lg - done >= 3 ? 3 : lg - done;
      radix(K, 1u << done, 0, z->tw_fwd.as<Fr>(), nullptr);
      done += K;
    }
    ZK_LAUNCH(k_join_abc, (size_t)n * B, 256, c->stream, abc, c->hsc.as<Fr>(), n, B);
  }
  CU(cudaGetLastError());
  return 0;
}

// ---- masked aggregation + model update on the device
extern "C" int zkfl_aggregate_updates(zkfl_ctx* c, const uint8_t* masked, const uint8_t* accept, uint32_t n_clients, uint32_t dim,
                                      double learning_rate, const double* model_in, uint8_t* agg_field_out, double* agg_mean_out,
                                      double* model_out, uint32_t* n_accepted) {
  if (!c || !masked || !model_in || !agg_mean_out || !model_out || n_clients == 0 || dim == 0) return fail(ZKFL_ERR_ARG, "bad argument");
  uint32_t count = 0;
  for (uint32_t i = 0; i < n_clients; i++) count += (!accept || accept[i]) ? 1u : 0u;
  if (n_accepted) *n_accepted = count;
  if (count == 0) return fail(ZKFL_ERR_ARG, "No verified updates to aggregate!");     // the reference returns null here
  CU(cudaSetDevice(c->device));
  const uint32_t n_chunks = (n_clients + ZK_AGG_CHUNK - 1) / ZK_AGG_CHUNK;
  DevBuf dMasked, dAccept, dPartial, dOut, dFlag;
  const size_t mbytes = (size_t)n_clients * dim * sizeof(Fr);
  TRY(dMasked.reserve(mbytes)); TRY(dPartial.reserve((size_t)dim * n_chunks * sizeof(Fr)));
  TRY(dOut.reserve((size_t)dim * (sizeof(Fr) + 3 * sizeof(double)))); TRY(dFlag.reserve(4));
  CU(cudaMemcpyAsync(dMasked.p, masked, mbytes, cudaMemcpyDefault, c->stream));          // host or device buffer
  const uint8_t* acc_dev = nullptr;
  if (accept) { TRY(dAccept.reserve(n_clients)); CU(cudaMemcpyAsync(dAccept.p, accept, n_clients, cudaMemcpyHostToDevice, c->stream)); acc_dev = dAccept.as<uint8_t>(); }
  CU(cudaMemsetAsync(dFlag.p, 0, 4, c->stream));
  Fr* agg = dOut.as<Fr>();
  double* d_model_in = (double*)(agg + dim);
  double* d_mean = d_model_in + dim;
  double* d_model_out = d_mean + dim;
  CU(cudaMemcpyAsync(d_model_in, model_in, dim * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  {
    Stage st(c, "aggregate");
    ZK_LAUNCH(k_agg_partial, (size_t)dim * n_chunks, 64, c->stream, dMasked.as<Fr>(), acc_dev, n_clients, dim, n_chunks, dPartial.as<Fr>(),
              dFlag.as<uint32_t>());
    ZK_LAUNCH(k_agg_final, dim, 32, c->stream, (const Fr*)dPartial.as<Fr>(), n_chunks, dim, count, learning_rate, (const double*)d_model_in, agg,
              d_mean, d_model_out);
    CU(cudaGetLastError());
  }
  uint32_t flag = 0;
  CU(cudaMemcpyAsync(&flag, dFlag.p, 4, cudaMemcpyDeviceToHost, c->stream));
  if (agg_field_out) CU(cudaMemcpyAsync(agg_field_out, agg, dim * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(agg_mean_out, d_mean, dim * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(model_out, d_model_out, dim * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (flag) return fail(ZKFL_ERR_ARG, "masked update not reduced mod r");
  return 0;
}
