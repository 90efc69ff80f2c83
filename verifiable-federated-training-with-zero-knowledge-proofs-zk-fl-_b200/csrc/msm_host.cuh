// Host side of the Pippenger pipeline, templated over the coordinate field; msm_g1.cu / msm_g2.cu instantiate it.
#pragma once
#include "host.h"
#include "k_msm.cuh"

static inline uint32_t affine_slots() { uint32_t K = env_u32("ZKFL_MSM_AFFINE_K", 64); return K < 1 ? 1 : (K > 4096 ? 4096 : K); }   // chunks per thread

// the operand-file form of the G2 accumulation (k_msm_accumulate_chunks_g2f): 9 Fq2 slots per thread in dynamic shared memory
static inline int msm_launch_g2f(zkfl_ctx* c, cudaStream_t stream, const void* bases, const uint32_t* sorted, const uint32_t* offsets,
                                 const uint32_t* counts, const MsmShape& s, uint32_t S, uint32_t cpr, void* buckets, void* head, void* tail) {
  const size_t total = (size_t)s.B * s.R * cpr;
  if (!total) return 0;
#ifndef ZKFL_EMUL
  const unsigned block = 128;
  const size_t smem = (size_t)G2F_SLOTS * 64 * block;
  if (!c->g2f_attr) { CU(cudaFuncSetAttribute(k_msm_accumulate_chunks_g2f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); c->g2f_attr = true; }
  k_msm_accumulate_chunks_g2f<<<(unsigned)((total + block - 1) / block), block, smem, stream>>>(
      (const G2Affine*)bases, sorted, offsets, counts, s, S, cpr, (G2Xyzz*)buckets, (G2Xyzz*)head, (G2Xyzz*)tail);
  zkrt::note_launch("k_msm_accumulate_chunks_g2f");
  if (zkrt::debug_sync()) zkrt::debug_check("k_msm_accumulate_chunks_g2f", stream);
#else
  ZK_LAUNCH(k_msm_accumulate_chunks_g2f, total, 128, stream, (const G2Affine*)bases, sorted, offsets, counts, s, S, cpr, (G2Xyzz*)buckets,
            (G2Xyzz*)head, (G2Xyzz*)tail);
#endif
  CU(cudaGetLastError());
  return 0;
}
// bucket accumulation of one MSM (slot = which of the five buffer sets) and the fix-up of the buckets cut by chunk borders, on the
// main stream; uses the lists left by msm_sort(gen).
template <class F>
int msm_accumulate(zkfl_ctx* c, const Affine<F>* bases, const MsmShape& s, int slot, const char* tag, int gen, cudaStream_t stream) {
  if (!stream) stream = c->stream;
  size_t rows = (size_t)s.B * s.R;
  TRY(c->buckets[slot].reserve(rows * s.nb * sizeof(Xyzz<F>)));
  const uint32_t S = s.lsS ? (1u << s.lsS) : accumulate_chunk(rows * s.cap), cpr = (s.cap + S - 1) / S;
  TRY(c->head[slot].reserve(rows * cpr * sizeof(Xyzz<F>)));
  TRY(c->tail[slot].reserve(rows * cpr * sizeof(Xyzz<F>)));
  const uint32_t* offsets = c->offsets[gen].as<uint32_t>();
  const uint32_t* counts = c->counts[gen].as<uint32_t>();
  Xyzz<F>* head = c->head[slot].as<Xyzz<F>>();
  Xyzz<F>* tail = c->tail[slot].as<Xyzz<F>>();
  if (s.lsS) {
    const uint32_t K = affine_slots();
    const size_t n_groups = rows * (cpr >> 5);
    if ((cpr >> 5) % K != 0 || rows >= 0xFFFFFFFFull) return fail(ZKFL_ERR_ARG, "batch-affine accumulation: bad list geometry");
    TRY(c->aff_acc.reserve(n_groups * 32 * sizeof(Affine<F>)));
    TRY(c->aff_pre.reserve(n_groups * 32 * sizeof(F)));
    Stage st(c, tag, stream);
    ZK_LAUNCH(k_msm_accumulate_affine<F>, n_groups / K * 32, 128, stream, bases, c->sorted[gen].as<uint32_t>(),
              c->skey.as<zk_key_t>(), offsets, counts, s, K, (uint32_t)rows,
              c->aff_acc.as<Affine<F>>(), c->aff_pre.as<F>(), c->buckets[slot].as<Xyzz<F>>(), head, tail);
  } else if (sizeof(F) > 32 && env_u32("ZKFL_G2_OPERAND_FILE", 0)) {   // opt-in: measured no faster, see k_msm.cuh
    Stage st(c, tag, stream);
    TRY(msm_launch_g2f(c, stream, (const void*)bases, c->sorted[gen].as<uint32_t>(), offsets, counts, s, S, cpr, c->buckets[slot].p, (void*)head, (void*)tail));
  } else {
    Stage st(c, tag, stream);
    ZK_LAUNCH(k_msm_accumulate_chunks<F>, rows * cpr, 128, stream, bases, c->sorted[gen].as<uint32_t>(), (const zk_key_t*)nullptr,
              offsets, counts, s, S, cpr, c->buckets[slot].as<Xyzz<F>>(), head, tail);
  }
  {
    Stage st(c, "msm_fixup", stream);
    // heavy buckets (runs over more than 16 chunks) go to warps, one per segment of 256 chunks; the queue lives in the context:
    // [slots used | heavy_cap records of four words], then heavy_cap segment sums.  Measured on B200 (2^20-domain proof / 2^20-point
    // MSM): span 16, segments of 256 -> fix-up 2.9 / 0.19 ms; span 4, segments of 64 (every bucket of such a proof spans ~7 chunks
    // and goes to a warp, whose five-round shuffle tree runs 32 lanes for ~8 partials) -> 8.6 / 2.0 ms.
    const uint32_t heavy_span = 16, heavy_cap = 16384;
    uint32_t* heavy = nullptr;
    Xyzz<F>* hsum = nullptr;
#ifndef ZKFL_EMUL
    if (env_u32("ZKFL_FIXUP_HEAVY", 1) && rows <= 64 && rows * s.nb < 0xFFFFFFFFull) {   // few rows: single proofs, split proofs
      const size_t qbytes = ((size_t)heavy_cap + 1) * 16;
      TRY(c->heavy[slot].reserve(qbytes + (size_t)heavy_cap * sizeof(Xyzz<Fq2>)));
      heavy = c->heavy[slot].as<uint32_t>();
      hsum = (Xyzz<F>*)((uint8_t*)c->heavy[slot].p + qbytes);
      CU(cudaMemsetAsync(heavy, 0, qbytes, stream));
    }
#endif
    // large batches: cut buckets are queued by how often they are cut and summed densely (k_msm_fixup_apply); at most one
    // bucket is cut per chunk border, so rows * cpr ids per queue suffice
    uint32_t* fq = nullptr;
    const size_t q_cap = rows * cpr;
    if (!heavy && rows >= env_u32("ZKFL_FIXUP_QUEUE_MIN_ROWS", 65) && rows * s.nb < 0xFFFFFFFFull && q_cap < 0xFFFFFFFFull && env_u32("ZKFL_FIXUP_QUEUE", 1)) {
      TRY(c->fixq.reserve((4 + 2 * q_cap) * 4));
      fq = c->fixq.as<uint32_t>();
      CU(cudaMemsetAsync(fq, 0, 16, stream));
    }
    const bool bound = env_u32("ZKFL_FIXUP_BOUND", 1) != 0;
    if (heavy)      // few rows: the latency form (inlined, overlapping products)
      ZK_LAUNCH((k_msm_fixup<F, 2>), rows * s.nb, 128, stream, offsets, counts, s, S, cpr, (const Xyzz<F>*)head, (const Xyzz<F>*)tail,
                c->buckets[slot].as<Xyzz<F>>(), heavy_span, heavy_cap, heavy, fq, (uint32_t)q_cap);
    else if (bound)
      ZK_LAUNCH((k_msm_fixup<F, 1>), rows * s.nb, 128, stream, offsets, counts, s, S, cpr, (const Xyzz<F>*)head, (const Xyzz<F>*)tail,
                c->buckets[slot].as<Xyzz<F>>(), heavy_span, heavy_cap, heavy, fq, (uint32_t)q_cap);
    else
      ZK_LAUNCH((k_msm_fixup<F, 0>), rows * s.nb, 128, stream, offsets, counts, s, S, cpr, (const Xyzz<F>*)head, (const Xyzz<F>*)tail,
                c->buckets[slot].as<Xyzz<F>>(), heavy_span, heavy_cap, heavy, fq, (uint32_t)q_cap);
    if (fq)
      for (uint32_t which = 0; which < 2; which++) {
        // the queue lengths are only known on the device: a grid for the worst case, unused threads exit at once
        if (bound)
          ZK_LAUNCH((k_msm_fixup_apply<F, 1>), q_cap, 128, stream, offsets, counts, s, S, cpr, (const Xyzz<F>*)head, (const Xyzz<F>*)tail,
                    c->buckets[slot].as<Xyzz<F>>(), (const uint32_t*)fq, (uint32_t)q_cap, which);
        else
          ZK_LAUNCH((k_msm_fixup_apply<F, 0>), q_cap, 128, stream, offsets, counts, s, S, cpr, (const Xyzz<F>*)head, (const Xyzz<F>*)tail,
                    c->buckets[slot].as<Xyzz<F>>(), (const uint32_t*)fq, (uint32_t)q_cap, which);
      }
#ifndef ZKFL_EMUL
    // the number of queued segments is only known on the device: a fixed grid of warps (unused slots exit at once) keeps the stream
    // free of host round trips; large batches never queue (their rows are small circuits) and skip the launches
    if (heavy) {
      ZK_LAUNCH(k_msm_fixup_heavy<F>, (size_t)heavy_cap * 32, 128, stream, offsets, counts, s, S, cpr, (const Xyzz<F>*)head, (const Xyzz<F>*)tail,
                c->buckets[slot].as<Xyzz<F>>(), heavy_cap, (const uint32_t*)heavy, hsum, 0);
      ZK_LAUNCH(k_msm_fixup_heavy<F>, (size_t)heavy_cap * 32, 128, stream, offsets, counts, s, S, cpr, (const Xyzz<F>*)head, (const Xyzz<F>*)tail,
                c->buckets[slot].as<Xyzz<F>>(), heavy_cap, (const uint32_t*)heavy, hsum, 1);
    }
#endif
  }
  CU(cudaGetLastError());
  return 0;
}
// bucket reduction sum_k (k+1) * B_k of one MSM on `stream` -> out[B]. Buffers must have been reserved (msm_reserve_reduce).
template <class F>
int msm_reduce(zkfl_ctx* c, const MsmShape& s, int slot, Xyzz<F>* out, cudaStream_t stream, const char* tag, int gen) {
  size_t rows = (size_t)s.B * s.R;
  ReducePlan p = reduce_plan(s);
  Xyzz<F>* R1 = c->Rs[slot].as<Xyzz<F>>();
  Xyzz<F>* T1 = c->Ts[slot].as<Xyzz<F>>();
  Xyzz<F>* R2 = c->lvl2[slot].as<Xyzz<F>>();
  Xyzz<F>* T2 = R2 + rows * p.N2;
  Xyzz<F>* RT = T2 + rows * p.N2;
  Stage st(c, tag, stream);
  if (reduce_deep(s)) {
    uint32_t lg = 0; while ((1u << lg) < s.nb) lg++;
    const Xyzz<F>* main_in = c->buckets[slot].as<Xyzz<F>>();
    uint32_t N = s.nb, n_pool = 0;
    int pp = 0;
    for (uint32_t done = 0; done < lg;) {
      const uint32_t lgL = lg - done >= 3 ? 3 : lg - done;
      Xyzz<F>* main_out = c->red_main[slot][pp].as<Xyzz<F>>();
      Xyzz<F>* pool_out = c->red_pool[slot][pp].as<Xyzz<F>>();
      const Xyzz<F>* pool_in = c->red_pool[slot][pp ^ 1].as<Xyzz<F>>();
      ZK_LAUNCH(k_reduce_bits_level<F>, rows * (N >> lgL) * (1 + n_pool + lgL), 64, stream, main_in, pool_in, n_pool, rows, N, lgL,
                main_out, pool_out);
      main_in = main_out; N >>= lgL; n_pool += lgL; done += lgL; pp ^= 1;
    }
    ZK_LAUNCH(k_reduce_bits_final<F>, rows, 32, stream, main_in, (const Xyzz<F>*)c->red_pool[slot][pp ^ 1].as<Xyzz<F>>(), n_pool, rows,
              c->win[slot].as<Xyzz<F>>());
    ZK_LAUNCH(k_msm_combine<F>, s.B, 32, stream, c->win[slot].as<Xyzz<F>>(), s, out);
    CU(cudaGetLastError());
    return 0;
  }
  if (env_u32("ZKFL_REDUCE_BOUND", 1)) {
    ZK_LAUNCH((k_reduce_level<F, 1>), rows * p.N1, 64, stream, (const Xyzz<F>*)c->buckets[slot].as<Xyzz<F>>(), rows, s.nb, p.L1, R1, T1);
    ZK_LAUNCH((k_reduce_level<F, 1>), rows * p.N2, 64, stream, (const Xyzz<F>*)R1, rows, p.N1, p.L2, R2, T2);
    ZK_LAUNCH((k_reduce_level<F, 1>), rows * p.N2, 64, stream, (const Xyzz<F>*)T1, rows, p.N1, p.L2, RT, (Xyzz<F>*)nullptr);
  } else {
    ZK_LAUNCH((k_reduce_level<F, 0>), rows * p.N1, 64, stream, (const Xyzz<F>*)c->buckets[slot].as<Xyzz<F>>(), rows, s.nb, p.L1, R1, T1);
    ZK_LAUNCH((k_reduce_level<F, 0>), rows * p.N2, 64, stream, (const Xyzz<F>*)R1, rows, p.N1, p.L2, R2, T2);
    ZK_LAUNCH((k_reduce_level<F, 0>), rows * p.N2, 64, stream, (const Xyzz<F>*)T1, rows, p.N1, p.L2, RT, (Xyzz<F>*)nullptr);
  }
  ZK_LAUNCH(k_reduce_final<F>, rows, 32, stream, (const Xyzz<F>*)R2, (const Xyzz<F>*)T2, (const Xyzz<F>*)RT, rows, p.N2, p.L1, p.L2,
            c->win[slot].as<Xyzz<F>>());
  ZK_LAUNCH(k_msm_combine<F>, s.B, 32, stream, c->win[slot].as<Xyzz<F>>(), s, out);
  CU(cudaGetLastError());
  return 0;
}
template <class F>
int msm_run(zkfl_ctx* c, const Affine<F>* bases, const MsmShape& s, Xyzz<F>* out, const char* acc_tag, const char* red_tag) {
  TRY(msm_accumulate<F>(c, bases, s, 0, acc_tag, 0));
  TRY(msm_reserve_reduce(c, s, 0, sizeof(Xyzz<F>)));
  return msm_reduce<F>(c, s, 0, out, c->stream, red_tag, 0);
}

template <class F>
int msm_precompute_windows(zkfl_ctx* c, const Affine<F>* raw, uint32_t cnt, uint32_t cw, uint32_t W, Affine<F>* table) {
  ZK_LAUNCH(k_precompute_windows<F>, cnt, 64, c->stream, raw, cnt, cw, W, table);
  CU(cudaGetLastError());
  return 0;
}
template <class F>
int msm_fixed_base_table(zkfl_ctx* c, const Affine<F>& base, Affine<F>* tab) {
  ZK_LAUNCH(k_fixed_base_table<F>, 32 * 256, 64, c->stream, base, tab);
  CU(cudaGetLastError());
  return 0;
}
template <class F>
int msm_to_affine_canonical(zkfl_ctx* c, const Xyzz<F>* in, size_t n, Affine<F>* out) {
  ZK_LAUNCH(k_to_affine_canonical<F>, n, n < 64 ? 32 : 64, c->stream, in, n, out);
  CU(cudaGetLastError());
  return 0;
}
template <class F>
int msm_sum_partials(zkfl_ctx* c, const Affine<F>* parts, uint32_t nparts, size_t part_stride, size_t n, Xyzz<F>* out) {
  ZK_LAUNCH(k_sum_partials<F>, n, 64, c->stream, parts, nparts, part_stride, (size_t)1, n, out);
  CU(cudaGetLastError());
  return 0;
}
template <class F>
int msm_gen_mul(zkfl_ctx* c, const Affine<F>& gen, const uint8_t* scalars, size_t n, uint8_t* out) {
  if (!c || !scalars || !out || n == 0) return fail(ZKFL_ERR_ARG, "bad argument");
  for (size_t i = 0; i < n; i++) if (!fr_bytes_lt_mod(scalars + 32 * i)) return fail(ZKFL_ERR_ARG, "scalar not reduced mod r");
  CU(cudaSetDevice(c->device));
  DevBuf sc, pts;
  TRY(sc.reserve(n * sizeof(Fr))); TRY(pts.reserve(n * sizeof(Affine<F>)));
  CU(cudaMemcpyAsync(sc.p, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  ZK_LAUNCH(k_gen_mul<F>, n, 64, c->stream, gen, sc.as<Fr>(), n, pts.as<Affine<F>>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, pts.p, n * sizeof(Affine<F>), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

template <class F>
int msm_point_scale(zkfl_ctx* c, const uint8_t* pts, const uint8_t* scalar, size_t n, uint8_t* out) {
  if (!c || !pts || !scalar || !out || n == 0) return fail(ZKFL_ERR_ARG, "bad argument");
  if (!fr_bytes_lt_mod(scalar)) return fail(ZKFL_ERR_ARG, "scalar not reduced mod r");
  CU(cudaSetDevice(c->device));
  DevBuf in, res;
  TRY(in.reserve(n * sizeof(Affine<F>))); TRY(res.reserve(n * sizeof(Affine<F>)));
  Fr k; memcpy(k.v, scalar, 32);
  CU(cudaMemcpyAsync(in.p, pts, n * sizeof(Affine<F>), cudaMemcpyHostToDevice, c->stream));
  ZK_LAUNCH(k_point_scale<F>, n, 64, c->stream, in.as<Affine<F>>(), k, n, res.as<Affine<F>>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, res.p, n * sizeof(Affine<F>), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

#define ZK_INSTANTIATE_MSM(F)                                                                                              \
  template int msm_accumulate<F>(zkfl_ctx*, const Affine<F>*, const MsmShape&, int, const char*, int, cudaStream_t);                       \
  template int msm_reduce<F>(zkfl_ctx*, const MsmShape&, int, Xyzz<F>*, cudaStream_t, const char*, int);                     \
  template int msm_run<F>(zkfl_ctx*, const Affine<F>*, const MsmShape&, Xyzz<F>*, const char*, const char*);               \
  template int msm_precompute_windows<F>(zkfl_ctx*, const Affine<F>*, uint32_t, uint32_t, uint32_t, Affine<F>*);           \
  template int msm_fixed_base_table<F>(zkfl_ctx*, const Affine<F>&, Affine<F>*);                                           \
  template int msm_to_affine_canonical<F>(zkfl_ctx*, const Xyzz<F>*, size_t, Affine<F>*);                                  \
  template int msm_sum_partials<F>(zkfl_ctx*, const Affine<F>*, uint32_t, size_t, size_t, Xyzz<F>*);                       \
  template int msm_gen_mul<F>(zkfl_ctx*, const Affine<F>&, const uint8_t*, size_t, uint8_t*);                              \
  template int msm_point_scale<F>(zkfl_ctx*, const uint8_t*, const uint8_t*, size_t, uint8_t*);
