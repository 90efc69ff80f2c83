// libzkfl.so, Pippenger pipeline: group-independent host parts (shape / cost model, bucket sort, reduction plan)
// and the G1 instantiation of the accumulate / reduce kernels.
#define ZK_K_MSM_SORT
#include "msm_host.cuh"

// sorted entries per accumulation thread.  Measured on B200 at 1024 sgd_verified proofs per step (proofs/s; accumulate G1 + G2 / fix-up
// ms): S = 24: 3 676 (192.0 / 29.2), 32: 3 700 (194.8 / 24.1), 40: 3 714, 48: 3 729, 64: 3 737 (200.8 / 15.0), 96: 3 698, 128: 3 668
// (208.4 / 12.3).  Short chunks make more of the additions free (the first entry of every run is a copy) but leave more partial sums
// to the fix-up, whose full additions cost 1.4 mixed ones at half the lane efficiency.
// By the size of the whole sort (`entries` = rows x list capacity): small jobs need the threads more than the cheaper fix-up -- a
// single sgd_verified proof (202 k entries) makes 6 k threads at 32 and runs 5.9 ms against 6.7 ms at 64 -- while from a 2^20-domain
// proof or a 2^20-point MSM on (15 M entries) 64 is at least as good (22.0 against 22.4 ms; 3.71 ms both ways).
uint32_t accumulate_chunk(size_t entries) { uint32_t S = env_u32("ZKFL_MSM_CHUNK", entries >= ((size_t)8 << 20) ? 64 : 32); return S < 4 ? 4 : S; }
// shared = all windows of a proof accumulate into ONE bucket set (bases table precomputed with the window shifts)
MsmShape msm_shape(uint32_t m, uint32_t B, bool shared, uint32_t force_c, uint32_t c_cap) {
  uint32_t best_c = 4; double best = 1e300;
  // one shared bucket set: up to 17 bits for proving keys (a split proof reduces the whole set on every rank, so it must stay small);
  // the standalone MSM over a resident table passes c_cap = 20 (2^19 buckets); per-window sets stop at 16
  const uint32_t c_max = c_cap ? c_cap : (shared ? 17 : 16);
  for (uint32_t c = 4; c <= c_max; c++) {
    double W = 254 / c + 1, nb = (double)(1u << (c - 1));
    double cost = shared ? W * (double)m + 2.6 * nb : W * ((double)m + 2.6 * nb);
    if (cost < best) { best = cost; best_c = c; }
  }
  uint32_t c = force_c ? force_c : env_u32(shared ? "ZKFL_MSM_C_SHARED" : "ZKFL_MSM_C", best_c);
  if (c < 2) c = 2;
  if (c > c_max) c = c_max;
  MsmShape s; s.m = m; s.B = B; s.c = c; s.W = 254 / c + 1; s.nb = 1u << (c - 1);
  s.R = shared ? 1 : s.W;
  s.cap = shared ? m * s.W : m;
  s.lsS = 0;
  // Batch-affine accumulation (k_msm_accumulate_affine): OPT-IN.  Measured on B200 at 1024 sgd_verified proofs it executes
  // fewer instructions per addition than the XYZZ chunk kernel (2390 vs ~2600) but runs at 22 % FMA-pipe utilisation
  // against 46 %: 130 registers, long-running warps (tail effect) and the serial latency of the shared inversion leave two
  // warps per scheduler on average (profiles/r01_ncu_full_k_msm_accumulate_affine.csv), so the XYZZ kernel stays the default.
  // ZKFL_MSM_AFFINE = 0 / unset: never, 1: always, 2: by size (needs K*S sorted entries per thread to fill the GPU).
  const uint32_t mode = env_u32("ZKFL_MSM_AFFINE", 0);
  uint32_t S = accumulate_chunk((size_t)B * s.R * s.cap), ls = 0;
  while ((1u << ls) < S) ls++;
  const double threads = (double)B * s.R * s.cap / ((double)(1u << ls) * affine_slots());
  if (mode == 1 || (mode != 0 && shared && threads >= 148.0 * 512.0)) {
    s.lsS = ls;
    const uint32_t unit = (32u << ls) * affine_slots();   // a warp owns K groups of 32 chunks of ONE row
    s.cap = (s.cap + unit - 1) / unit * unit;
  }
  return s;
}
// three-level reduction tree over nb = L1 * L2 * N2 buckets
ReducePlan reduce_plan(const MsmShape& s) {
  uint32_t lg = 0; while ((1u << lg) < s.nb) lg++;
  uint32_t l1 = (lg + 2) / 3, l2 = (lg - l1 + 1) / 2;
  ReducePlan p; p.L1 = 1u << l1; p.L2 = 1u << l2; p.N1 = s.nb >> l1; p.N2 = p.N1 >> l2;
  return p;
}

int msm_sort_reserve(zkfl_ctx* c, const MsmShape& s, int gen) {
  size_t rows = (size_t)s.B * s.R;
  uint32_t nchunk = (s.nb + ZK_SCAN_CHUNK - 1) / ZK_SCAN_CHUNK;
  TRY(c->counts[gen].reserve(rows * s.nb * 4));
  TRY(c->offsets[gen].reserve(rows * s.nb * 4));
  TRY(c->cursors.reserve(rows * s.nb * 4));
  TRY(c->chunk_sums.reserve(rows * nchunk * 4));
  TRY(c->sorted[gen].reserve(rows * s.cap * 4));
  if (s.lsS != 0) TRY(c->skey.reserve(rows * s.cap * sizeof(zk_key_t)));   // only the batch-affine accumulation walks the list by keys
  return 0;
}
int msm_sort(zkfl_ctx* c, const Fr* scalars, const uint8_t* skip, const MsmShape& s, int gen, cudaStream_t stream) {
  if (!stream) stream = c->stream;
  size_t rows = (size_t)s.B * s.R;
  uint32_t nchunk = (s.nb + ZK_SCAN_CHUNK - 1) / ZK_SCAN_CHUNK;
  TRY(msm_sort_reserve(c, s, gen));
  zk_key_t* keys = s.lsS != 0 ? c->skey.as<zk_key_t>() : nullptr;
#ifndef ZKFL_EMUL
  // OPT-IN (ZKFL_MSM_SORT_CTA=1): one CTA per proof with the histogram in shared memory.  Measured on B200 at 1024 proofs: 8.6 ms per
  // sort against ~7 ms for the global-atomics passes below -- ~300 proofs are in flight at once, their 1.2 MB list regions no longer fit
  // the 126 MB L2 together and the scattered 4 / 2-byte list writes become DRAM read-modify-writes; the proof-major thread order of
  // the passes below keeps ~30 proofs in flight, L2-resident.
  if (s.R == 1 && rows >= 32 && s.nb <= 32768 && env_u32("ZKFL_MSM_SORT_CTA", 0)) {
    const size_t smem = ((size_t)s.nb + 32) * 4;
    if (!c->sort_attr) { CU(cudaFuncSetAttribute(k_msm_sort_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (32768 + 32) * 4)); c->sort_attr = true; }
    k_msm_sort_cta<<<(unsigned)rows, 1024, smem, stream>>>(scalars, skip, s, c->offsets[gen].as<uint32_t>(), c->counts[gen].as<uint32_t>(),
                                                             c->sorted[gen].as<uint32_t>(), keys);
    zkrt::note_launch("k_msm_sort_cta");
    if (zkrt::debug_sync()) zkrt::debug_check("k_msm_sort_cta", stream);
    CU(cudaGetLastError());
    return 0;
  }
#endif
  CU(cudaMemsetAsync(c->counts[gen].p, 0, rows * s.nb * 4, stream));
  ZK_LAUNCH(k_msm_count, (size_t)s.m * s.B, 256, stream, scalars, skip, s, c->counts[gen].as<uint32_t>());
  ZK_LAUNCH(k_msm_scan_chunks, rows * nchunk, 128, stream, c->counts[gen].as<uint32_t>(), s, c->chunk_sums.as<uint32_t>());
  ZK_LAUNCH(k_msm_scan_write, rows * nchunk, 128, stream, c->counts[gen].as<uint32_t>(), c->chunk_sums.as<uint32_t>(), s,
            c->offsets[gen].as<uint32_t>(), c->cursors.as<uint32_t>());
  ZK_LAUNCH(k_msm_scatter, (size_t)s.m * s.B, 256, stream, scalars, skip, s, c->cursors.as<uint32_t>(), c->sorted[gen].as<uint32_t>(),
            keys);
  CU(cudaGetLastError());
  return 0;
}
// few rows: the bit-decomposed tree of plain sums (k_reduce_bits_level) instead of the three-level running sums
bool reduce_deep(const MsmShape& s) {
  const uint32_t mode = env_u32("ZKFL_REDUCE_DEEP", 2);   // 0 never, 1 always, 2 by size
  return mode == 1 || (mode == 2 && (size_t)s.B * s.R <= 32 && s.nb >= 64);
}
int msm_reserve_reduce(zkfl_ctx* c, const MsmShape& s, int slot, size_t elem) {
  size_t rows = (size_t)s.B * s.R;
  ReducePlan p = reduce_plan(s);
  if (reduce_deep(s)) {
    for (int k = 0; k < 2; k++) {
      TRY(c->red_main[slot][k].reserve(rows * (s.nb / 2) * elem));
      TRY(c->red_pool[slot][k].reserve(rows * s.nb * elem));   // <= 3/8 + 6/64 + ... of nb per row, with slack for small fan-ins
    }
  }
  TRY(c->Rs[slot].reserve(rows * p.N1 * elem));
  TRY(c->Ts[slot].reserve(rows * p.N1 * elem));
  TRY(c->lvl2[slot].reserve(3 * rows * p.N2 * elem));
  TRY(c->win[slot].reserve(rows * elem));
  return 0;
}
int msm_range_mask(zkfl_ctx* c, const uint8_t* base_skip, uint32_t m, uint32_t lo, uint32_t hi, uint8_t* out) {
  ZK_LAUNCH(k_range_mask, m, 256, c->stream, base_skip, m, lo, hi, out);
  CU(cudaGetLastError());
  return 0;
}

ZK_INSTANTIATE_MSM(Fq)
