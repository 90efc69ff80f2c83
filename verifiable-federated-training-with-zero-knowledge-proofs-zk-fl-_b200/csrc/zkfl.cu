// libzkfl.so host side: artefact parsing (.zkey / .r1cs / .zkwp), HBM residency, orchestration of the proving
// pipeline on one CUDA stream per context (side streams for the bucket reductions), and the C ABI declared in
// include/zkfl.h.  Kernel families live in witness.cu, msm_g1.cu, msm_g2.cu, verify.cu (see host.h).
#include "host.h"
#include "k_fin.cuh"
#include "k_bench.cuh"

#ifdef ZKFL_EMUL
thread_local zk_emul_idx zk_emul_cur;
#endif

// ------------------------------------------------------------------------------------ errors / counters
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches(0);
namespace zkrt {
void note_launch(const char* name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  static int trace = -1;   // ZKFL_TRACE_LAUNCHES=1: kernel names on stderr (tests use it to see which path ran)
  if (trace < 0) { const char* v = getenv("ZKFL_TRACE_LAUNCHES"); trace = (v && *v && *v != '0') ? 1 : 0; }
  if (trace == 1) fprintf(stderr, "[zkfl launch] %s\n", name);
}
bool debug_sync() {
  static int on = -1;
  if (on < 0) { const char* v = getenv("ZKFL_DEBUG_SYNC"); on = (v && *v && *v != '0') ? 1 : 0; }
  return on == 1;
}
void debug_check(const char* name, cudaStream_t stream) {
#ifndef ZKFL_EMUL
  cudaError_t e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  fprintf(stderr, "[zkfl] %-40s %s\n", name, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
#else
  (void)name; (void)stream;
#endif
}
}  // namespace zkrt
int zk_fail(int code, const std::string& msg) { g_err = msg; return code; }
uint64_t zk_launches_now() { return g_launches.load(); }

// ------------------------------------------------------------------------------------ host field helpers
static Fr fr_from_bytes_canonical(const uint8_t* p) { Fr r; memcpy(r.v, p, 32); return r; }
static Fr fr_pow(Fr base, const uint32_t* e, int nwords) {
  Fr r = Fr::one();
  for (int i = nwords * 32 - 1; i >= 0; i--) { r = r.sqr(); if ((e[i >> 5] >> (i & 31)) & 1) r = r * base; }
  return r;
}
Fr fr_root_of_unity(int power) {  // ffjavascript: nqr = 5, w[28] = 5^((r-1)/2^28), w[k] = w[k+1]^2
  uint32_t e[8];
  for (int i = 0; i < 8; i++) e[i] = FrP::mod(i);
  e[0] -= 1;
  for (int i = 0; i < 8; i++) e[i] = (e[i] >> 28) | (i < 7 ? e[i + 1] << 4 : 0);
  Fr five = Fr::zero(); five.v[0] = 5;
  Fr w = fr_pow(five.to_mont(), e, 8);
  for (int i = 28; i > power; i--) w = w.sqr();
  return w;
}
bool fr_bytes_lt_mod(const uint8_t* p) {
  uint32_t v[8]; memcpy(v, p, 32);
  for (int i = 7; i >= 0; i--) { if (v[i] < FrP::mod(i)) return true; if (v[i] > FrP::mod(i)) return false; }
  return false;
}

// ------------------------------------------------------------------------------------ iden3 binfile
int parse_sections(const uint8_t* d, size_t len, const char* magic, std::map<uint32_t, Sec>& out) {
  if (!d || len < 12 || memcmp(d, magic, 4)) return fail(ZKFL_ERR_FORMAT, std::string("bad magic, expected ") + magic);
  uint32_t n; memcpy(&n, d + 8, 4);
  size_t p = 12;
  for (uint32_t i = 0; i < n; i++) {
    if (p + 12 > len) return fail(ZKFL_ERR_FORMAT, "truncated section table");
    uint32_t id; uint64_t l; memcpy(&id, d + p, 4); memcpy(&l, d + p + 4, 8); p += 12;
    if (l > len - p) return fail(ZKFL_ERR_FORMAT, "truncated section");
    if (!out.count(id)) out[id] = Sec{d + p, l};
    p += l;
  }
  return 0;
}

// builds CSR from COO triples (row, wire, coef) with counting sort by row
static void coo_to_csr(uint32_t n_rows, const std::vector<uint32_t>& rows, const std::vector<uint32_t>& wires,
                       const std::vector<Fr>& coefs, CsrHost& out) {
  out.row_off.assign(n_rows + 1, 0);
  for (uint32_t r : rows) out.row_off[r + 1]++;
  for (uint32_t i = 0; i < n_rows; i++) out.row_off[i + 1] += out.row_off[i];
  out.wire.resize(rows.size()); out.coef.resize(rows.size());
  std::vector<uint32_t> cur(out.row_off.begin(), out.row_off.end() - 1);
  for (size_t i = 0; i < rows.size(); i++) { uint32_t p = cur[rows[i]]++; out.wire[p] = wires[i]; out.coef[p] = coefs[i]; }
}

// ------------------------------------------------------------------------------------ prove pipeline (witness in c->w)
static int finalize_from_sums(zkfl_ctx* c, const zkfl_zkey* z, const Fr* rs_dev, uint32_t B);
// part / nparts: this context handles the point range [part*m/nparts, (part+1)*m/nparts) of every MSM (nparts == 1: all).
// finalize == false stops after the five MSM sums (res_g1 / res_g2).
static int prove_from_device_witness(zkfl_ctx* c, const zkfl_zkey* z, const Fr* rs_dev, uint32_t B, uint32_t part = 0,
                                     uint32_t nparts = 1, bool finalize = true, const zkfl_r1cs* check = nullptr) {
  const uint32_t n = z->domain, m = z->n_vars;
  Fr* w = c->w.as<Fr>();
  TRY(c->res_g1.reserve(4 * (size_t)B * sizeof(G1Xyzz)));
  TRY(c->res_g2.reserve((size_t)B * sizeof(G2Xyzz)));
  G1Xyzz* r1 = c->res_g1.as<G1Xyzz>();
  G2Xyzz* r2 = c->res_g2.as<G2Xyzz>();
  MsmShape sw = msm_shape(m, B, true, z->c_w);
  MsmShape sh = msm_shape(n, B, true, z->c_h);
  // slots: 0 = A, 1 = C, 2 = B1, 3 = H (G1), 4 = B2 (G2). All allocations happen before any side-stream work
  // (cudaMalloc / cudaFree inside reserve() would serialise the device).
  for (int slot = 0; slot < 3; slot++) TRY(msm_reserve_reduce(c, sw, slot, sizeof(G1Xyzz)));
  TRY(msm_reserve_reduce(c, sh, 3, sizeof(G1Xyzz)));
  TRY(msm_reserve_reduce(c, sw, 4, sizeof(G2Xyzz)));
  TRY(msm_sort_reserve(c, sw, 0)); TRY(msm_sort_reserve(c, sw, 1)); TRY(msm_sort_reserve(c, sh, 2));
  if (!c->side[0]) {
    for (int i = 0; i < 5; i++) {
      CU(cudaStreamCreate(&c->side[i]));
      CU(cudaEventCreate(&c->ev_acc[i]));
      CU(cudaEventCreate(&c->ev_red[i]));
    }
    CU(cudaStreamCreate(&c->sort_stream));
    CU(cudaEventCreate(&c->ev_in)); CU(cudaEventCreate(&c->ev_hsc));
    for (int i = 0; i < 3; i++) CU(cudaEventCreate(&c->ev_sort[i]));
  }
  const uint8_t *skip_w = nullptr, *skip_wb = z->skipB.as<uint8_t>(), *skip_h = nullptr;
  if (nparts > 1) {
    TRY(c->mask_w.reserve(m)); TRY(c->mask_wb.reserve(m)); TRY(c->mask_h.reserve(n));
    uint32_t lo = (uint32_t)((uint64_t)m * part / nparts), hi = (uint32_t)((uint64_t)m * (part + 1) / nparts);
    uint32_t hlo = (uint32_t)((uint64_t)n * part / nparts), hhi = (uint32_t)((uint64_t)n * (part + 1) / nparts);
    TRY(msm_range_mask(c, nullptr, m, lo, hi, c->mask_w.as<uint8_t>()));
    TRY(msm_range_mask(c, z->skipB.as<uint8_t>(), m, lo, hi, c->mask_wb.as<uint8_t>()));
    TRY(msm_range_mask(c, nullptr, n, hlo, hhi, c->mask_h.as<uint8_t>()));
    skip_w = c->mask_w.as<uint8_t>(); skip_wb = c->mask_wb.as<uint8_t>(); skip_h = c->mask_h.as<uint8_t>();
  }
  // FEW ROWS (single proofs, the per-rank share of a split proof; ZKFL_MSM_CONCURRENT=0 turns it off): no kernel fills the GPU and
  // everything is latency -- the five MSMs run side by side, each whole (accumulate, fix-up, reduce) on its own stream: the witness
  // sorts come first (they need only the witness), A, C, B1, B2 start behind them, and A.w / B.w, the NTTs and the H sort run on the
  // main stream meanwhile; H follows on its stream.  The heavy-bucket queues are per MSM for that (zkfl_ctx::heavy[slot]).
  const bool concurrent = B <= 64 && sw.lsS == 0 && sh.lsS == 0 && env_u32("ZKFL_MSM_CONCURRENT", 1) != 0;
  // OPT-IN (ZKFL_SORT_OVERLAP=1, large batches): the sorts (bound by L2 atomics and scattered stores) on their own stream beside the
  // product-bound stages -- the two witness sorts beside A.w / B.w and the NTTs, the H sort beside the accumulation of the witness
  // MSMs; each sort has its own list / counts / offsets set for that.  MEASURED AND REJECTED as the default on B200 (1024
  // sgd_verified proofs per step): 3 796 proofs/s against 3 822 in the serial order -- the accumulation already fills every SM
  // (4 CTAs x 128 registers), so the sort CTAs only take residency from it and every co-running stage stretches (NTT 15.1 -> 20.5 ms,
  // fix-up 9.3 -> 18.6 ms, accumulation 130 -> 133 ms, sorts 19 -> 58 ms of stream time).
  const bool overlap = !concurrent && sw.lsS == 0 && sh.lsS == 0 && env_u32("ZKFL_SORT_OVERLAP", 0) != 0;
  cudaStream_t ss = overlap ? c->sort_stream : c->stream;
  auto sort = [&](const char* tag, const Fr* sc, const uint8_t* skip, const MsmShape& shp, int gen) -> int {
    { Stage st(c, tag, ss); TRY(msm_sort(c, sc, skip, shp, gen, ss)); }
    if (overlap || concurrent) CU(cudaEventRecord(c->ev_sort[gen], ss));
    return 0;
  };
  auto sorted_ready = [&](int gen) -> int { if (overlap) CU(cudaStreamWaitEvent(c->stream, c->ev_sort[gen], 0)); return 0; };
  // one MSM: serial order -- accumulate + fix-up on the main stream, the latency-bound reduction on the slot's side stream;
  // concurrent -- all three on the side stream, behind the sort of its lists
  auto msm_g1 = [&](int slot, int gen, const G1Affine* bases, const MsmShape& shp, G1Xyzz* out) -> int {
    cudaStream_t as = concurrent ? c->side[slot] : c->stream;
    if (concurrent) CU(cudaStreamWaitEvent(as, c->ev_sort[gen], 0));
    TRY(msm_accumulate<Fq>(c, bases, shp, slot, "msm_acc_g1", gen, as));
    if (!concurrent) { CU(cudaEventRecord(c->ev_acc[slot], c->stream)); CU(cudaStreamWaitEvent(c->side[slot], c->ev_acc[slot], 0)); }
    TRY(msm_reduce<Fq>(c, shp, slot, out, c->side[slot], "msm_reduce_g1", gen));
    CU(cudaEventRecord(c->ev_red[slot], c->side[slot]));
    return 0;
  };
  auto msm_g2 = [&](int slot, int gen, const G2Affine* bases, const MsmShape& shp, G2Xyzz* out) -> int {
    cudaStream_t as = concurrent ? c->side[slot] : c->stream;
    if (concurrent) CU(cudaStreamWaitEvent(as, c->ev_sort[gen], 0));
    TRY(msm_accumulate<Fq2>(c, bases, shp, slot, "msm_acc_g2", gen, as));
    if (!concurrent) { CU(cudaEventRecord(c->ev_acc[slot], c->stream)); CU(cudaStreamWaitEvent(c->side[slot], c->ev_acc[slot], 0)); }
    TRY(msm_reduce<Fq2>(c, shp, slot, out, c->side[slot], "msm_reduce_g2", gen));
    CU(cudaEventRecord(c->ev_red[slot], c->side[slot]));
    return 0;
  };
  if (concurrent) {
    TRY(sort("msm_sort_w", w, skip_w, sw, 0));
    TRY(sort("msm_sort_w", w, skip_wb, sw, 1));
    TRY(msm_g1(0, 0, z->pA.as<G1Affine>(), sw, r1));
    TRY(msm_g1(1, 0, z->pC.as<G1Affine>(), sw, r1 + 2 * (size_t)B));
    TRY(msm_g2(4, 1, z->pB2.as<G2Affine>(), sw, r2));          // the longest of the four first
    TRY(msm_g1(2, 1, z->pB1.as<G1Affine>(), sw, r1 + B));
    TRY(run_h_poly(c, z, B, check));
    TRY(sort("msm_sort_h", c->hsc.as<Fr>(), skip_h, sh, 2));
    TRY(msm_g1(3, 2, z->pH.as<G1Affine>(), sh, r1 + 3 * (size_t)B));
  } else {
    if (overlap) {
      CU(cudaEventRecord(c->ev_in, c->stream));            // witness (and range masks) complete; earlier passes' accumulations too
      CU(cudaStreamWaitEvent(ss, c->ev_in, 0));
      TRY(sort("msm_sort_w", w, skip_w, sw, 0));
      TRY(sort("msm_sort_w", w, skip_wb, sw, 1));
    }
    TRY(run_h_poly(c, z, B, check));
    if (overlap) {
      CU(cudaEventRecord(c->ev_hsc, c->stream));
      CU(cudaStreamWaitEvent(ss, c->ev_hsc, 0));
      TRY(sort("msm_sort_h", c->hsc.as<Fr>(), skip_h, sh, 2));
    }
    if (!overlap) TRY(sort("msm_sort_w", w, skip_w, sw, 0));
    TRY(sorted_ready(0));
    TRY(msm_g1(0, 0, z->pA.as<G1Affine>(), sw, r1));
    TRY(msm_g1(1, 0, z->pC.as<G1Affine>(), sw, r1 + 2 * (size_t)B));
    if (!overlap) TRY(sort("msm_sort_w", w, skip_wb, sw, 1));
    TRY(sorted_ready(1));
    TRY(msm_g1(2, 1, z->pB1.as<G1Affine>(), sw, r1 + B));
    TRY(msm_g2(4, 1, z->pB2.as<G2Affine>(), sw, r2));
    if (!overlap) TRY(sort("msm_sort_h", c->hsc.as<Fr>(), skip_h, sh, 2));
    TRY(sorted_ready(2));
    TRY(msm_g1(3, 2, z->pH.as<G1Affine>(), sh, r1 + 3 * (size_t)B));
  }
  for (int i = 0; i < 5; i++) CU(cudaStreamWaitEvent(c->stream, c->ev_red[i], 0));
  if (!finalize) return 0;
  return finalize_from_sums(c, z, rs_dev, B);
}

// blinding / assembly from the five MSM sums in res_g1 ([4][B]: A, B1, C, H) and res_g2 ([B]: B2)
static int finalize_from_sums(zkfl_ctx* c, const zkfl_zkey* z, const Fr* rs_dev, uint32_t B) {
  G1Xyzz* r1 = c->res_g1.as<G1Xyzz>();
  {
    Stage st(c, "finalize");
    TRY(c->t_g1.reserve(3 * (size_t)B * sizeof(G1Xyzz)));
    TRY(c->t_g2.reserve((size_t)B * sizeof(G2Xyzz)));
    TRY(c->pis.reserve(2 * (size_t)B * sizeof(G1Xyzz)));
    TRY(c->var.reserve(2 * (size_t)B * sizeof(G1Xyzz)));
    TRY(c->proofs.reserve((size_t)B * 256));
    ZK_LAUNCH(k_fin_fixed, (size_t)B * 4, 32, c->stream, z->tab_d1.as<G1Affine>(), z->tab_d2.as<G2Affine>(), rs_dev, B,
              c->t_g1.as<G1Xyzz>(), c->t_g2.as<G2Xyzz>());
    ZK_LAUNCH(k_fin_var, (size_t)B * 2, 32, c->stream, z->vk, rs_dev, B, r1, c->t_g1.as<G1Xyzz>(), c->pis.as<G1Xyzz>(),
              c->var.as<G1Xyzz>());
    ZK_LAUNCH(k_fin_write, (size_t)B * 3, 32, c->stream, z->vk, B, r1, c->res_g2.as<G2Xyzz>(), c->t_g1.as<G1Xyzz>(),
              c->t_g2.as<G2Xyzz>(), c->pis.as<G1Xyzz>(), c->var.as<G1Xyzz>(), c->proofs.as<Fq>());
  }
  CU(cudaGetLastError());
  return 0;
}

static int stage_rs(zkfl_ctx* c, const uint8_t* rs, int B) {
  std::vector<uint8_t> tmp;
  if (!rs) {  // snarkjs: r, s <- Fr.random()
    tmp.resize((size_t)B * 64);
    FILE* f = fopen("/dev/urandom", "rb");
    if (!f) return fail(ZKFL_ERR_ARG, "cannot open /dev/urandom");
    for (size_t i = 0; i < (size_t)B * 2; i++) {
      do {
        if (fread(&tmp[32 * i], 1, 32, f) != 32) { fclose(f); return fail(ZKFL_ERR_ARG, "urandom read failed"); }
        tmp[32 * i + 31] &= 0x3f;
      } while (!fr_bytes_lt_mod(&tmp[32 * i]));
    }
    fclose(f);
    rs = tmp.data();
  } else {
    for (size_t i = 0; i < (size_t)B * 2; i++)
      if (!fr_bytes_lt_mod(rs + 32 * i)) return fail(ZKFL_ERR_ARG, "blinding scalar not reduced mod r");
  }
  TRY(c->stage_rs.reserve((size_t)B * 64));
  CU(cudaMemcpyAsync(c->stage_rs.p, rs, (size_t)B * 64, cudaMemcpyHostToDevice, c->stream));
  if (!tmp.empty()) CU(cudaStreamSynchronize(c->stream));
  return 0;
}

static int fetch_publics(zkfl_ctx* c, uint32_t n_public, uint32_t B, uint8_t* publics_out) {
  if (!publics_out || !n_public) return 0;
  TRY(c->pubs.reserve((size_t)n_public * B * sizeof(Fr)));
  TRY(zk_soa_to_aos(c, c->w.as<Fr>() + B, c->pubs.as<Fr>(), n_public, B));
  CU(cudaMemcpyAsync(publics_out, c->pubs.p, (size_t)n_public * B * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  return 0;
}


// ------------------------------------------------------------------------------------ C ABI
extern "C" {

const char* zkfl_last_error(void) { return g_err.c_str(); }
const char* zkfl_version(void) {
#ifdef ZKFL_EMUL
  return "zkfl 0.1.0 (host emulation, tests only)";
#else
  return "zkfl 0.1.0 (sm_100a)";
#endif
}
uint64_t zkfl_launch_count(void) { return g_launches.load(); }

int zkfl_ctx_create(int device, zkfl_ctx** out) {
  if (!out) return fail(ZKFL_ERR_ARG, "out is NULL");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return fail(ZKFL_ERR_CUDA, "no CUDA device available (libzkfl has no CPU path)");
  if (device < 0 || device >= n) return fail(ZKFL_ERR_ARG, "device index out of range");
  CU(cudaSetDevice(device));
#ifndef ZKFL_EMUL
  {  // the group-law helpers are real calls (noinline): give their frames room
    const char* v = getenv("ZKFL_STACK_BYTES");
    size_t want = v && *v ? (size_t)strtoul(v, nullptr, 10) : 0;
    if (want) CU(cudaDeviceSetLimit(cudaLimitStackSize, want));
  }
#endif
  zkfl_ctx* c = new zkfl_ctx();
  c->device = device;
  cudaError_t e = cudaStreamCreate(&c->stream);
  if (e != cudaSuccess) { delete c; return fail(ZKFL_ERR_CUDA, "cudaStreamCreate failed"); }
  *out = c;
  return 0;
}
void zkfl_ctx_free(zkfl_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& r : c->pending) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  if (c->t0) { cudaEventDestroy(c->t0); cudaEventDestroy(c->t1); }
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->chk_host) cudaFreeHost(c->chk_host);
  for (int i = 0; i < 5; i++) {
    if (!c->side[i]) continue;
    if (i == 0 && c->sort_stream) {
      cudaStreamSynchronize(c->sort_stream);
      cudaStreamDestroy(c->sort_stream);
      cudaEventDestroy(c->ev_in); cudaEventDestroy(c->ev_hsc);
      for (int k = 0; k < 3; k++) cudaEventDestroy(c->ev_sort[k]);
    }
    cudaStreamSynchronize(c->side[i]);
    cudaEventDestroy(c->ev_acc[i]);
    cudaEventDestroy(c->ev_red[i]);
    cudaStreamDestroy(c->side[i]);
  }
  cudaStreamDestroy(c->stream);
  delete c;
}

int zkfl_prof_enable(zkfl_ctx* c, int on) {
  if (!c) return fail(ZKFL_ERR_ARG, "ctx is NULL");
  c->prof = on != 0;
  if (on) { c->agg.clear(); c->order.clear(); }
  return 0;
}
int zkfl_prof_read(zkfl_ctx* c, char* buf, size_t cap) {
  if (!c || !buf || !cap) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaStreamSynchronize(c->stream));
  for (auto& r : c->pending) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (!c->agg.count(r.name)) c->order.push_back(r.name);
    ProfAgg& a = c->agg[r.name];
    a.ms += ms; a.launches += r.launches; a.calls++;
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  c->pending.clear();
  std::string s;
  for (auto& name : c->order) {
    const ProfAgg& a = c->agg[name];
    char line[256];
    snprintf(line, sizeof line, "%s %.6f %llu %llu\n", name.c_str(), a.ms, (unsigned long long)a.launches, (unsigned long long)a.calls);
    s += line;
  }
  if (s.size() + 1 > cap) return fail(ZKFL_ERR_ARG, "buffer too small");
  memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}

// ---- circuit (.zkwp)
int zkfl_circuit_load(zkfl_ctx* c, const uint8_t* d, size_t len, zkfl_circuit** out) {
  if (!c || !out) return fail(ZKFL_ERR_ARG, "bad argument");
  std::map<uint32_t, Sec> S;
  TRY(parse_sections(d, len, "zkwp", S));
  for (uint32_t id = 1; id <= 7; id++) if (!S.count(id)) return fail(ZKFL_ERR_FORMAT, "zkwp: missing section");
  if (!S.count(9) || S[9].len < 8 || S[9].len % 4) return fail(ZKFL_ERR_FORMAT, "zkwp: missing level schedule");
  if (S[1].len != 32) return fail(ZKFL_ERR_FORMAT, "zkwp: bad header");
  uint32_t h[8]; memcpy(h, S[1].p, 32);
  uint32_t n_lcs = h[4], n_terms = h[5], n_pos = h[6], n_widths = h[7];
  if (S[2].len != 20ull * h[3] || S[3].len != 4ull * (n_lcs + 1) || S[4].len != 4ull * n_terms || S[5].len != 32ull * n_terms ||
      S[6].len != 4ull * n_pos)
    return fail(ZKFL_ERR_FORMAT, "zkwp: section size mismatch");
  std::unique_ptr<zkfl_circuit> k(new zkfl_circuit());
  k->ctx = c; k->n_wires = h[0]; k->n_public = h[1]; k->n_inputs = h[2]; k->n_ops = h[3];
  CU(cudaSetDevice(c->device));
  TRY(upload(c, k->ops, (const uint32_t*)S[2].p, 5 * (size_t)h[3]));
  TRY(upload(c, k->lc_off, (const uint32_t*)S[3].p, (size_t)n_lcs + 1));
  TRY(upload(c, k->lc_wire, (const uint32_t*)S[4].p, n_terms));
  std::vector<Fr> coef(n_terms);
  for (uint32_t i = 0; i < n_terms; i++) {
    if (!fr_bytes_lt_mod(S[5].p + 32 * (size_t)i)) return fail(ZKFL_ERR_FORMAT, "zkwp: coefficient not reduced");
    coef[i] = fr_from_bytes_canonical(S[5].p + 32 * (size_t)i).to_mont().to_mont();  // coef * R^2
  }
  TRY(upload(c, k->lc_coef, coef.data(), coef.size()));
  TRY(upload(c, k->pos_in, (const uint32_t*)S[6].p, n_pos));
  // poseidon constants -> Montgomery, one pool
  std::vector<Fr> pool;
  struct Pk { uint32_t t, rounds, rp; size_t c_off, m_off; };
  std::vector<Pk> pks;
  const uint8_t* q = S[7].p; const uint8_t* qe = q + S[7].len;
  for (uint32_t w = 0; w < n_widths; w++) {
    if (q + 16 > qe) return fail(ZKFL_ERR_FORMAT, "zkwp: truncated poseidon table");
    uint32_t hh[4]; memcpy(hh, q, 16); q += 16;
    Pk p; p.t = hh[0]; p.rounds = hh[1]; p.rp = hh[2];
    if (p.t < 2 || p.t > 17 || p.rounds != 8 + p.rp) return fail(ZKFL_ERR_FORMAT, "zkwp: bad poseidon width");
    size_t cnt = (size_t)p.rounds * p.t + (size_t)p.t * p.t;
    if (q + 32 * cnt > qe) return fail(ZKFL_ERR_FORMAT, "zkwp: truncated poseidon constants");
    p.c_off = pool.size(); p.m_off = p.c_off + (size_t)p.rounds * p.t;
    for (size_t i = 0; i < cnt; i++) { pool.push_back(fr_from_bytes_canonical(q).to_mont()); q += 32; }
    pks.push_back(p);
  }
  TRY(upload(c, k->pconst, pool.data(), pool.size()));
  memset(&k->dev, 0, sizeof k->dev);
  k->dev.n_wires = k->n_wires; k->dev.n_inputs = k->n_inputs; k->dev.n_ops = k->n_ops;
  k->dev.ops = k->ops.as<uint32_t>(); k->dev.lc_off = k->lc_off.as<uint32_t>(); k->dev.lc_wire = k->lc_wire.as<uint32_t>();
  k->dev.lc_coef = k->lc_coef.as<Fr>(); k->dev.pos_in = k->pos_in.as<uint32_t>();
  for (auto& p : pks) {
    k->dev.pk[p.t].rounds = p.rounds; k->dev.pk[p.t].rp = p.rp;
    k->dev.pk[p.t].C = k->pconst.as<Fr>() + p.c_off; k->dev.pk[p.t].M = k->pconst.as<Fr>() + p.m_off;
  }
  k->level_off.resize(S[9].len / 4);
  memcpy(k->level_off.data(), S[9].p, S[9].len);
  if (k->level_off.front() != 0 || k->level_off.back() != k->n_ops) return fail(ZKFL_ERR_FORMAT, "zkwp: bad level schedule");
  for (size_t i = 0; i + 1 < k->level_off.size(); i++)
    if (k->level_off[i] > k->level_off[i + 1]) return fail(ZKFL_ERR_FORMAT, "zkwp: bad level schedule");
  // A .zkwp comes from disk (it stands where circom's .wasm stood): every index and extent is checked here, so a malformed
  // program is rejected at load time and can never make a kernel read or write outside its buffers.
  if ((uint64_t)k->n_inputs + 1 > k->n_wires || k->n_public > k->n_inputs) return fail(ZKFL_ERR_FORMAT, "zkwp: bad header counts");
  const uint32_t* lc_off = (const uint32_t*)S[3].p;
  const uint32_t* lc_wire = (const uint32_t*)S[4].p;
  const uint32_t* pos_in = (const uint32_t*)S[6].p;
  if (lc_off[0] != 0 || lc_off[n_lcs] != n_terms) return fail(ZKFL_ERR_FORMAT, "zkwp: bad linear-combination offsets");
  for (uint32_t i = 0; i < n_lcs; i++) if (lc_off[i] > lc_off[i + 1]) return fail(ZKFL_ERR_FORMAT, "zkwp: bad linear-combination offsets");
  for (uint32_t i = 0; i < n_terms; i++) if (lc_wire[i] >= k->n_wires) return fail(ZKFL_ERR_FORMAT, "zkwp: term wire out of range");
  for (uint32_t i = 0; i < n_pos; i++) if (pos_in[i] >= k->n_wires) return fail(ZKFL_ERR_FORMAT, "zkwp: poseidon input wire out of range");
  const uint32_t* ops = (const uint32_t*)S[2].p;
  for (uint32_t o = 0; o < k->n_ops; o++) {
    const uint32_t* op = ops + 5 * (size_t)o;
    const uint32_t code = op[0], dst = op[1];
    uint64_t n_out = 1;
    if (code < 1 || code > 4) return fail(ZKFL_ERR_FORMAT, "zkwp: bad op code");
    if (code <= 3 && op[2] >= n_lcs) return fail(ZKFL_ERR_FORMAT, "zkwp: linear combination out of range");
    if (code == 2 && (op[3] >= n_lcs || (op[4] != 0xFFFFFFFFu && op[4] >= n_lcs))) return fail(ZKFL_ERR_FORMAT, "zkwp: linear combination out of range");
    if (code == 3) { if (op[3] > 254) return fail(ZKFL_ERR_FORMAT, "zkwp: bit decomposition wider than the field"); n_out = op[3]; }
    if (code == 4) {
      const uint32_t t = op[2];
      if (t < 2 || t > 17 || !k->dev.pk[t].C) return fail(ZKFL_ERR_FORMAT, "zkwp: poseidon width without constants");
      if ((uint64_t)op[3] + (t - 1) > n_pos) return fail(ZKFL_ERR_FORMAT, "zkwp: poseidon input list out of range");
      n_out = 3ull * (8ull * t + k->dev.pk[t].rp) + 1;   // x^2, x^4, x^5 per S-box, then the output
    }
    if (dst <= k->n_inputs || (uint64_t)dst + n_out > k->n_wires) return fail(ZKFL_ERR_FORMAT, "zkwp: op output out of range");
  }
  *out = k.release();
  return 0;
}
void zkfl_circuit_free(zkfl_circuit* k) { if (k) { cudaSetDevice(k->ctx->device); delete k; } }
int zkfl_circuit_info(const zkfl_circuit* k, uint32_t info[4]) {
  if (!k || !info) return fail(ZKFL_ERR_ARG, "bad argument");
  info[0] = k->n_wires; info[1] = k->n_public; info[2] = k->n_inputs; info[3] = k->n_ops;
  return 0;
}

// ---- r1cs
int zkfl_r1cs_load(zkfl_ctx* c, const uint8_t* d, size_t len, zkfl_r1cs** out) {
  if (!c || !out) return fail(ZKFL_ERR_ARG, "bad argument");
  std::map<uint32_t, Sec> S;
  TRY(parse_sections(d, len, "r1cs", S));
  if (!S.count(1) || !S.count(2) || S[1].len < 64) return fail(ZKFL_ERR_FORMAT, "r1cs: missing section");
  const uint8_t* h = S[1].p;
  uint32_t n8; memcpy(&n8, h, 4);
  if (n8 != 32) return fail(ZKFL_ERR_FORMAT, "r1cs: field size");
  uint32_t n_wires, n_constraints; memcpy(&n_wires, h + 36, 4); memcpy(&n_constraints, h + 60, 4);
  std::vector<uint32_t> rows[3], wires[3]; std::vector<Fr> coefs[3];
  const uint8_t* p = S[2].p; const uint8_t* pe = p + S[2].len;
  for (uint32_t r = 0; r < n_constraints; r++)
    for (int k = 0; k < 3; k++) {
      if (p + 4 > pe) return fail(ZKFL_ERR_FORMAT, "r1cs: truncated");
      uint32_t nt; memcpy(&nt, p, 4); p += 4;
      if ((uint64_t)(pe - p) < 36ull * nt) return fail(ZKFL_ERR_FORMAT, "r1cs: truncated");
      for (uint32_t t = 0; t < nt; t++) {
        uint32_t wi; memcpy(&wi, p, 4);
        if (wi >= n_wires) return fail(ZKFL_ERR_FORMAT, "r1cs: wire out of range");
        rows[k].push_back(r); wires[k].push_back(wi);
        coefs[k].push_back(fr_from_bytes_canonical(p + 4).to_mont().to_mont());
        p += 36;
      }
    }
  std::unique_ptr<zkfl_r1cs> r(new zkfl_r1cs());
  r->ctx = c; r->n_wires = n_wires; r->n_constraints = n_constraints;
  CU(cudaSetDevice(c->device));
  CsrHost hA, hB, hC;
  coo_to_csr(n_constraints, rows[0], wires[0], coefs[0], hA);
  coo_to_csr(n_constraints, rows[1], wires[1], coefs[1], hB);
  coo_to_csr(n_constraints, rows[2], wires[2], coefs[2], hC);
  TRY(upload_csr(c, hA, r->A)); TRY(upload_csr(c, hB, r->B)); TRY(upload_csr(c, hC, r->C));
  *out = r.release();
  return 0;
}
void zkfl_r1cs_free(zkfl_r1cs* r) { if (r) { cudaSetDevice(r->ctx->device); delete r; } }

// ---- zkey
// nparts > 1: the key will prove ONE proof split over nparts GPUs (zkfl_groth16_msm_partials): every rank accumulates 1 / nparts
// of the points but reduces a whole bucket set, so the window tables are sized for the rank's share (fewer, smaller bucket sets)
static int zkey_load_impl(zkfl_ctx* c, const uint8_t* d, size_t len, uint32_t nparts, zkfl_zkey** out) {
  if (!c || !out || nparts == 0) return fail(ZKFL_ERR_ARG, "bad argument");
  std::map<uint32_t, Sec> S;
  TRY(parse_sections(d, len, "zkey", S));
  for (uint32_t id : {1u, 2u, 4u, 5u, 6u, 7u, 8u, 9u}) if (!S.count(id)) return fail(ZKFL_ERR_FORMAT, "zkey: missing section");
  uint32_t proto; if (S[1].len < 4) return fail(ZKFL_ERR_FORMAT, "zkey: bad section 1");
  memcpy(&proto, S[1].p, 4);
  if (proto != 1) return fail(ZKFL_ERR_FORMAT, "zkey: not a groth16 key");
  const uint8_t* h = S[2].p;
  if (S[2].len != 84 + 64 * 3 + 128 * 3) return fail(ZKFL_ERR_FORMAT, "zkey: bad header size");
  uint32_t n8q, n8r; memcpy(&n8q, h, 4); memcpy(&n8r, h + 36, 4);
  if (n8q != 32 || n8r != 32) return fail(ZKFL_ERR_FORMAT, "zkey: not BN254");
  for (int i = 0; i < 8; i++) {
    uint32_t a, b; memcpy(&a, h + 4 + 4 * i, 4); memcpy(&b, h + 40 + 4 * i, 4);
    if (a != FqP::mod(i) || b != FrP::mod(i)) return fail(ZKFL_ERR_FORMAT, "zkey: not BN254");
  }
  std::unique_ptr<zkfl_zkey> z(new zkfl_zkey());
  z->ctx = c;
  memcpy(&z->n_vars, h + 72, 4); memcpy(&z->n_public, h + 76, 4); memcpy(&z->domain, h + 80, 4);
  uint32_t n = z->domain, m = z->n_vars, l = z->n_public;
  if (n < 2 || (n & (n - 1)) || m <= l) return fail(ZKFL_ERR_FORMAT, "zkey: bad sizes");
  z->log_n = 0; while ((1u << z->log_n) < n) z->log_n++;
  if (z->log_n > 27) return fail(ZKFL_ERR_FORMAT, "zkey: domain too large");
  size_t p = 84;
  memcpy(&z->vk.alpha1, h + p, 64); p += 64;
  memcpy(&z->vk.beta1, h + p, 64); p += 64;
  memcpy(&z->vk.beta2, h + p, 128); p += 128;
  p += 128;
  memcpy(&z->vk.delta1, h + p, 64); p += 64;
  memcpy(&z->vk.delta2, h + p, 128);
  if (S[5].len != 64ull * m || S[6].len != 64ull * m || S[7].len != 128ull * m || S[8].len != 64ull * (m - l - 1) || S[9].len != 64ull * n)
    return fail(ZKFL_ERR_FORMAT, "zkey: point section size mismatch");
  uint32_t n_coef; if (S[4].len < 4) return fail(ZKFL_ERR_FORMAT, "zkey: bad coeffs");
  memcpy(&n_coef, S[4].p, 4);
  if (S[4].len != 4 + 44ull * n_coef) return fail(ZKFL_ERR_FORMAT, "zkey: coeff section size mismatch");
  std::vector<uint32_t> rows[2], wires[2]; std::vector<Fr> coefs[2];
  for (uint32_t i = 0; i < n_coef; i++) {
    const uint8_t* e = S[4].p + 4 + 44 * (size_t)i;
    uint32_t mt, cc, ss; memcpy(&mt, e, 4); memcpy(&cc, e + 4, 4); memcpy(&ss, e + 8, 4);
    if (mt > 1 || cc >= n || ss >= m) return fail(ZKFL_ERR_FORMAT, "zkey: coefficient out of range");
    rows[mt].push_back(cc); wires[mt].push_back(ss); coefs[mt].push_back(fr_from_bytes_canonical(e + 12));  // already coef * R^2
  }
  CU(cudaSetDevice(c->device));
  CsrHost hA, hB;
  coo_to_csr(n, rows[0], wires[0], coefs[0], hA);
  coo_to_csr(n, rows[1], wires[1], coefs[1], hB);
  TRY(upload_csr(c, hA, z->A)); TRY(upload_csr(c, hB, z->B));
  // bases -> window-shifted tables 2^(c*j) * P_i (built on the device once per key)
  z->c_w = msm_shape(m / nparts + 1, 1, true, env_u32("ZKFL_MSM_C_W", 0)).c;   // tuning knobs; 0 = cost model
  z->c_h = msm_shape(n / nparts + 1, 1, true, env_u32("ZKFL_MSM_C_H", 0)).c;
  auto build_table = [&](const uint8_t* pts, size_t bytes, uint32_t cnt, uint32_t cw, bool g2, DevBuf& out) -> int {
    DevBuf raw;
    TRY(upload(c, raw, pts, bytes));
    uint32_t W = 254 / cw + 1;
    TRY(out.reserve((size_t)W * bytes));
    if (g2) TRY(msm_precompute_windows<Fq2>(c, raw.as<G2Affine>(), cnt, cw, W, out.as<G2Affine>()));
    else TRY(msm_precompute_windows<Fq>(c, raw.as<G1Affine>(), cnt, cw, W, out.as<G1Affine>()));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
  };
  TRY(build_table(S[5].p, S[5].len, m, z->c_w, false, z->pA));
  TRY(build_table(S[6].p, S[6].len, m, z->c_w, false, z->pB1));
  TRY(build_table(S[7].p, S[7].len, m, z->c_w, true, z->pB2));
  TRY(build_table(S[9].p, S[9].len, n, z->c_h, false, z->pH));
  TRY(z->tab_d1.reserve(32 * 256 * sizeof(G1Affine)));
  TRY(z->tab_d2.reserve(32 * 256 * sizeof(G2Affine)));
  TRY(msm_fixed_base_table<Fq>(c, z->vk.delta1, z->tab_d1.as<G1Affine>()));
  TRY(msm_fixed_base_table<Fq2>(c, z->vk.delta2, z->tab_d2.as<G2Affine>()));
  CU(cudaStreamSynchronize(c->stream));
  {  // wires without a B-query point (absent from the B matrix): dropped when sorting for the B1 / B2 MSMs
    std::vector<uint8_t> skip(m, 0);
    static const uint8_t zero64[64] = {0};
    for (uint32_t i = 0; i < m; i++) skip[i] = memcmp(S[6].p + 64 * (size_t)i, zero64, 64) == 0;
    TRY(upload(c, z->skipB, skip.data(), skip.size()));
  }
  {  // C bases padded to n_vars so all four witness MSMs share one sorted index list
    std::vector<uint8_t> full(64 * (size_t)m, 0);
    memcpy(full.data() + 64 * (size_t)(l + 1), S[8].p, S[8].len);
    TRY(build_table(full.data(), full.size(), m, z->c_w, false, z->pC));
  }
  {  // twiddles: w^k, w^-k (k < n/2); coset[p] = n^-1 * inc^bitrev(p), inc = w_{2n}
    Fr wn = fr_root_of_unity((int)z->log_n), wi = wn.inv();
    std::vector<Fr> f(n / 2), iv(n / 2), cs(n);
    Fr a = Fr::one(), b = Fr::one();
    for (uint32_t k = 0; k < n / 2; k++) { f[k] = a; iv[k] = b; a = a * wn; b = b * wi; }
    Fr inc = fr_root_of_unity((int)z->log_n + 1);
    Fr nn = Fr::zero(); nn.v[0] = n;
    Fr cur = nn.to_mont().inv();
    for (uint32_t i = 0; i < n; i++) {
      uint32_t rev = 0;
      for (uint32_t bit = 0; bit < z->log_n; bit++) if (i & (1u << bit)) rev |= 1u << (z->log_n - 1 - bit);
      cs[rev] = cur;  // position rev holds coefficient i after the DIF pass
      cur = cur * inc;
    }
    TRY(upload(c, z->tw_fwd, f.data(), f.size())); TRY(upload(c, z->tw_inv, iv.data(), iv.size()));
    TRY(upload(c, z->coset, cs.data(), cs.size()));
  }
  *out = z.release();
  return 0;
}
int zkfl_zkey_load(zkfl_ctx* c, const uint8_t* d, size_t len, zkfl_zkey** out) { return zkey_load_impl(c, d, len, 1, out); }
int zkfl_zkey_load_split(zkfl_ctx* c, const uint8_t* d, size_t len, uint32_t nparts, zkfl_zkey** out) { return zkey_load_impl(c, d, len, nparts, out); }
void zkfl_zkey_free(zkfl_zkey* z) { if (z) { cudaSetDevice(z->ctx->device); delete z; } }
int zkfl_zkey_info(const zkfl_zkey* z, uint32_t info[3]) {
  if (!z || !info) return fail(ZKFL_ERR_ARG, "bad argument");
  info[0] = z->n_vars; info[1] = z->n_public; info[2] = z->domain;
  return 0;
}

// ---- witness
int zkfl_wtns_calculate_batch(zkfl_ctx* c, const zkfl_circuit* k, const zkfl_r1cs* r, const uint8_t* inputs, int B,
                              uint8_t* wtns_out, uint32_t* first_bad) {
  if (!c || !k || !inputs || B <= 0) return fail(ZKFL_ERR_ARG, "bad argument");
  if (r && r->n_wires != k->n_wires) return fail(ZKFL_ERR_ARG, "r1cs does not match circuit");
  for (size_t i = 0; i < (size_t)B * k->n_inputs; i++)
    if (!fr_bytes_lt_mod(inputs + 32 * i)) return fail(ZKFL_ERR_ARG, "input not reduced mod r");
  CU(cudaSetDevice(c->device));
  checks_reset(c);
  TRY(run_witness(c, k, inputs, (uint32_t)B));
  if (wtns_out) {
    TRY(c->aos.reserve((size_t)k->n_wires * B * sizeof(Fr)));
    TRY(zk_soa_to_aos(c, c->w.as<Fr>(), c->aos.as<Fr>(), k->n_wires, (uint32_t)B));
    CU(cudaMemcpyAsync(wtns_out, c->aos.p, (size_t)k->n_wires * B * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  }
  int rc = 0;
  if (r) rc = check_r1cs_device(c, r, (uint32_t)B, first_bad);
  CU(cudaStreamSynchronize(c->stream));
  return rc;
}
// runs the program and returns only the selected wires (B x n_sel field elements); the witness stays in HBM
int zkfl_wtns_eval_wires(zkfl_ctx* c, const zkfl_circuit* k, const uint8_t* inputs, int B, const uint32_t* wires, uint32_t n_sel,
                         uint8_t* out) {
  if (!c || !k || !inputs || !wires || !out || B <= 0 || n_sel == 0) return fail(ZKFL_ERR_ARG, "bad argument");
  for (uint32_t i = 0; i < n_sel; i++) if (wires[i] >= k->n_wires) return fail(ZKFL_ERR_ARG, "wire index out of range");
  for (size_t i = 0; i < (size_t)B * k->n_inputs; i++)
    if (!fr_bytes_lt_mod(inputs + 32 * i)) return fail(ZKFL_ERR_ARG, "input not reduced mod r");
  CU(cudaSetDevice(c->device));
  checks_reset(c);
  TRY(run_witness(c, k, inputs, (uint32_t)B));
  TRY(c->bad.reserve((size_t)n_sel * 4));
  TRY(c->aos.reserve((size_t)n_sel * B * sizeof(Fr)));
  CU(cudaMemcpyAsync(c->bad.p, wires, (size_t)n_sel * 4, cudaMemcpyHostToDevice, c->stream));
  TRY(zk_gather_wires(c, c->w.as<Fr>(), c->bad.as<uint32_t>(), n_sel, (uint32_t)B, c->aos.as<Fr>()));
  CU(cudaMemcpyAsync(out, c->aos.p, (size_t)n_sel * B * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
int zkfl_r1cs_check_batch(zkfl_ctx* c, const zkfl_r1cs* r, const uint8_t* wtns, int B, uint32_t* first_bad) {
  if (!c || !r || !wtns || B <= 0) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  checks_reset(c);
  size_t cnt = (size_t)r->n_wires * B;
  TRY(c->aos.reserve(cnt * sizeof(Fr))); TRY(c->w.reserve(cnt * sizeof(Fr)));
  CU(cudaMemcpyAsync(c->aos.p, wtns, cnt * sizeof(Fr), cudaMemcpyDefault, c->stream));   // pageable, pinned or device memory
  TRY(zk_aos_to_soa(c, c->aos.as<Fr>(), c->w.as<Fr>(), r->n_wires, (uint32_t)B, 0u));
  c->w_wires = r->n_wires; c->w_B = (uint32_t)B;
  return check_r1cs_device(c, r, (uint32_t)B, first_bad);
}

// ---- prove
int zkfl_groth16_prove_batch(zkfl_ctx* c, const zkfl_zkey* z, const uint8_t* wtns, const uint8_t* rs, int B,
                             uint8_t* proofs_out, uint8_t* publics_out) {
  if (!c || !z || !wtns || !proofs_out || B <= 0) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  checks_reset(c);
  size_t cnt = (size_t)z->n_vars * B;
  TRY(stage_rs(c, rs, B));
  {
    Stage st(c, "upload_wtns");
    TRY(c->aos.reserve(cnt * sizeof(Fr))); TRY(c->w.reserve(cnt * sizeof(Fr)));
    CU(cudaMemcpyAsync(c->aos.p, wtns, cnt * sizeof(Fr), cudaMemcpyDefault, c->stream));   // pageable, pinned or device memory
    TRY(zk_aos_to_soa(c, c->aos.as<Fr>(), c->w.as<Fr>(), z->n_vars, (uint32_t)B, 0u));
    c->w_wires = z->n_vars; c->w_B = (uint32_t)B;
    TRY(check_wtns_launch(c, z->n_vars, (uint32_t)B));   // snarkjs reads the .wtns through its field class: reduced values, w[0] = 1
  }
  TRY(prove_from_device_witness(c, z, c->stage_rs.as<Fr>(), (uint32_t)B));
  CU(cudaMemcpyAsync(proofs_out, c->proofs.p, (size_t)B * 256, cudaMemcpyDeviceToHost, c->stream));
  TRY(fetch_publics(c, z->n_public, (uint32_t)B, publics_out));
  CU(cudaStreamSynchronize(c->stream));
  return checks_result(c, nullptr, 0);
}
static int validate_inputs(const uint8_t* inputs, size_t count) {
  for (size_t i = 0; i < count; i++)
    if (!fr_bytes_lt_mod(inputs + 32 * i)) return fail(ZKFL_ERR_ARG, "input not reduced mod r");
  return 0;
}
// circom's witness calculator aborts at the first failing `===`, so `fullProve` never proves an unsatisfied witness.  With
// r1cs != NULL the constraint check runs on the HBM-resident witness inside the same stream-ordered pass (no second witness
// run, no copy of the witness to the host); its verdict is read after the final synchronisation: any violation -> the proofs
// are NOT returned (buffers zeroed), ZKFL_ERR_ASSERT, first_bad[b] = first violated row or 0xFFFFFFFF.
int zkfl_groth16_full_prove_batch(zkfl_ctx* c, const zkfl_circuit* k, const zkfl_zkey* z, const zkfl_r1cs* r, const uint8_t* inputs,
                                  const uint8_t* rs, int B, uint8_t* proofs_out, uint8_t* publics_out, uint32_t* first_bad) {
  if (!c || !k || !z || !inputs || !proofs_out || B <= 0) return fail(ZKFL_ERR_ARG, "bad argument");
  if (k->n_wires != z->n_vars || k->n_public != z->n_public) return fail(ZKFL_ERR_ARG, "circuit and zkey do not match");
  if (r && r->n_wires != k->n_wires) return fail(ZKFL_ERR_ARG, "r1cs does not match circuit");
  TRY(validate_inputs(inputs, (size_t)B * k->n_inputs));
  CU(cudaSetDevice(c->device));
  checks_reset(c);
  TRY(stage_rs(c, rs, B));
  TRY(run_witness(c, k, inputs, (uint32_t)B));
  TRY(prove_from_device_witness(c, z, c->stage_rs.as<Fr>(), (uint32_t)B, 0, 1, true, r));
  CU(cudaMemcpyAsync(proofs_out, c->proofs.p, (size_t)B * 256, cudaMemcpyDeviceToHost, c->stream));
  TRY(fetch_publics(c, z->n_public, (uint32_t)B, publics_out));
  CU(cudaStreamSynchronize(c->stream));
  int rc = checks_result(c, first_bad, (uint32_t)B);
  if (rc) memset(proofs_out, 0, (size_t)B * 256);
  return rc;
}
int zkfl_full_prove_stage(zkfl_ctx* c, const zkfl_circuit* k, const zkfl_zkey* z, const uint8_t* inputs, const uint8_t* rs, int B) {
  if (!c || !k || !z || !inputs || B <= 0) return fail(ZKFL_ERR_ARG, "bad argument");
  if (k->n_wires != z->n_vars) return fail(ZKFL_ERR_ARG, "circuit and zkey do not match");
  TRY(validate_inputs(inputs, (size_t)B * k->n_inputs));
  CU(cudaSetDevice(c->device));
  TRY(stage_rs(c, rs, B));
  TRY(c->stage_in.reserve((size_t)k->n_inputs * B * sizeof(Fr)));
  CU(cudaMemcpyAsync(c->stage_in.p, inputs, (size_t)k->n_inputs * B * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
int zkfl_full_prove_run(zkfl_ctx* c, const zkfl_circuit* k, const zkfl_zkey* z, const zkfl_r1cs* r, int B) {
  if (!c || !k || !z || B <= 0) return fail(ZKFL_ERR_ARG, "bad argument");
  if (r && r->n_wires != k->n_wires) return fail(ZKFL_ERR_ARG, "r1cs does not match circuit");
  CU(cudaSetDevice(c->device));
  checks_reset(c);
  TRY(run_witness(c, k, nullptr, (uint32_t)B));
  TRY(prove_from_device_witness(c, z, c->stage_rs.as<Fr>(), (uint32_t)B, 0, 1, true, r));
  return 0;  // asynchronous: the caller brackets with its own events / zkfl_full_prove_fetch (which reports the check)
}
int zkfl_full_prove_fetch(zkfl_ctx* c, int B, uint8_t* proofs_out, uint32_t* first_bad) {
  if (!c || B <= 0) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  if (proofs_out) CU(cudaMemcpyAsync(proofs_out, c->proofs.p, (size_t)B * 256, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  int rc = checks_result(c, first_bad, (uint32_t)B);
  if (rc && proofs_out) memset(proofs_out, 0, (size_t)B * 256);
  return rc;
}

// ---- single large proof split over several GPUs: per-rank MSM partials, then gather + add + blind
int zkfl_groth16_msm_partials(zkfl_ctx* c, const zkfl_zkey* z, const uint8_t* wtns, int B, uint32_t part, uint32_t nparts,
                              uint8_t* partials_out) {
  if (!c || !z || !partials_out || B <= 0 || nparts == 0 || part >= nparts) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  checks_reset(c);
  size_t cnt = (size_t)z->n_vars * B;
  if (wtns) {
    TRY(c->aos.reserve(cnt * sizeof(Fr))); TRY(c->w.reserve(cnt * sizeof(Fr)));
    CU(cudaMemcpyAsync(c->aos.p, wtns, cnt * sizeof(Fr), cudaMemcpyDefault, c->stream));   // pageable, pinned or device memory
    TRY(zk_aos_to_soa(c, c->aos.as<Fr>(), c->w.as<Fr>(), z->n_vars, (uint32_t)B, 0u));
    c->w_wires = z->n_vars; c->w_B = (uint32_t)B;
    TRY(check_wtns_launch(c, z->n_vars, (uint32_t)B));
  } else if (c->w.cap < cnt * sizeof(Fr) || c->w_wires != z->n_vars || c->w_B != (uint32_t)B) {
    // wtns == NULL: the witness the last witness calculation of this context left in HBM (no upload, no host copy)
    return fail(ZKFL_ERR_ARG, "no resident witness of this shape: run zkfl_wtns_calculate_batch on this context first");
  }
  TRY(prove_from_device_witness(c, z, nullptr, (uint32_t)B, part, nparts, false));
  // layout per proof b: A | B1 | C | H (64 B each, affine canonical) | B2 (128 B)  -> stored as [5 blocks][B]
  TRY(c->part_out.reserve((size_t)B * 384));
  uint8_t* o = c->part_out.as<uint8_t>();
  TRY(msm_to_affine_canonical<Fq>(c, c->res_g1.as<G1Xyzz>(), (size_t)4 * B, (G1Affine*)o));
  TRY(msm_to_affine_canonical<Fq2>(c, c->res_g2.as<G2Xyzz>(), (size_t)B, (G2Affine*)(o + (size_t)B * 256)));
  CU(cudaMemcpyAsync(partials_out, o, (size_t)B * 384, cudaMemcpyDefault, c->stream));   // host or device buffer (NCCL exchanges device memory)
  CU(cudaStreamSynchronize(c->stream));
  return checks_result(c, nullptr, 0);
}
int zkfl_groth16_finalize(zkfl_ctx* c, const zkfl_zkey* z, const uint8_t* partials, uint32_t nparts, const uint8_t* rs, int B,
                          uint8_t* proofs_out) {
  if (!c || !z || !partials || !proofs_out || B <= 0 || nparts == 0) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  TRY(stage_rs(c, rs, B));
  size_t per = (size_t)B * 384;
  TRY(c->part_in.reserve(per * nparts));
  TRY(c->res_g1.reserve(4 * (size_t)B * sizeof(G1Xyzz)));
  TRY(c->res_g2.reserve((size_t)B * sizeof(G2Xyzz)));
  CU(cudaMemcpyAsync(c->part_in.p, partials, per * nparts, cudaMemcpyDefault, c->stream));   // host or device buffer
  const uint8_t* in = c->part_in.as<uint8_t>();
  TRY(msm_sum_partials<Fq>(c, (const G1Affine*)in, nparts, per / sizeof(G1Affine), (size_t)4 * B, c->res_g1.as<G1Xyzz>()));
  TRY(msm_sum_partials<Fq2>(c, (const G2Affine*)(in + (size_t)B * 256), nparts, per / sizeof(G2Affine), (size_t)B, c->res_g2.as<G2Xyzz>()));
  TRY(finalize_from_sums(c, z, c->stage_rs.as<Fr>(), (uint32_t)B));
  CU(cudaMemcpyAsync(proofs_out, c->proofs.p, (size_t)B * 256, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- standalone MSM
static int msm_bases_upload(zkfl_ctx* c, const uint8_t* bases, size_t n, int group, void** handle) {
  if (!c || !bases || !handle || (group != 1 && group != 2) || n == 0 || n > 0x7FFFFFFFu) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  std::unique_ptr<MsmBases> b(new MsmBases());
  b->ctx = c; b->group = group; b->n = n;
  TRY(upload(c, b->pts, bases, n * (group == 1 ? 64 : 128)));
  *handle = b.release();
  return 0;
}
// the window-shifted table of a resident base set (zkfl_msm_bases_load builds it; the one-shot calls do not)
static int msm_bases_build_table(zkfl_ctx* c, MsmBases* b) {
  // window bits of the table: measured on B200 at 2^20 points (tests/dev/msm_c_sweep.py): c = 16 / 17 / 19 / 20 -> 4.11 / 3.88 / 4.36 /
  // 4.11 ms -- beyond 17 bits the single-row sort and the 3 additions per bucket of the latency reduction cost more than the saved
  // windows, so the cost model is capped at 17; ZKFL_MSM_C_TABLE forces any width up to 20
  const uint32_t c_forced = env_u32("ZKFL_MSM_C_TABLE", 0);
  const uint32_t cw = msm_shape((uint32_t)b->n, 1, true, c_forced, c_forced ? 20 : 17).c, W = 254 / cw + 1;
  TRY(b->table.reserve((size_t)W * b->n * (b->group == 1 ? 64 : 128)));
  if (b->group == 1) TRY(msm_precompute_windows<Fq>(c, b->pts.as<G1Affine>(), (uint32_t)b->n, cw, W, b->table.as<G1Affine>()));
  else TRY(msm_precompute_windows<Fq2>(c, b->pts.as<G2Affine>(), (uint32_t)b->n, cw, W, b->table.as<G2Affine>()));
  CU(cudaStreamSynchronize(c->stream));
  b->c_tab = cw;
  return 0;
}
int zkfl_msm_bases_load(zkfl_ctx* c, const uint8_t* bases, size_t n, int group, void** handle) {
  TRY(msm_bases_upload(c, bases, n, group, handle));
  if (n >= 1024 && env_u32("ZKFL_MSM_TABLE", 1)) {
    int rc = msm_bases_build_table(c, (MsmBases*)*handle);
    if (rc) { zkfl_msm_bases_free(*handle); *handle = nullptr; return rc; }
  }
  return 0;
}
void zkfl_msm_bases_free(void* h) { if (h) { MsmBases* b = (MsmBases*)h; cudaSetDevice(b->ctx->device); delete b; } }
int zkfl_msm_run(zkfl_ctx* c, void* handle, const uint8_t* scalars, size_t n, uint8_t* out) {
  MsmBases* b = (MsmBases*)handle;
  if (!c || !b || n == 0 || n > b->n) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  TRY(c->msm_sc.reserve(n * sizeof(Fr)));
  if (scalars) CU(cudaMemcpyAsync(c->msm_sc.p, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  const bool tab = b->c_tab && n == b->n;      // the table is indexed j * n + i: whole base set only
  MsmShape s = tab ? msm_shape((uint32_t)n, 1, true, b->c_tab, 20) : msm_shape((uint32_t)n, 1, false);
  { Stage st(c, "msm_sort"); TRY(msm_sort(c, c->msm_sc.as<Fr>(), nullptr, s)); }
  TRY(c->msm_out.reserve(sizeof(G2Xyzz) + sizeof(G2Affine)));
  uint8_t* o = c->msm_out.as<uint8_t>();
  if (b->group == 1) {
    TRY(msm_run<Fq>(c, tab ? b->table.as<G1Affine>() : b->pts.as<G1Affine>(), s, (G1Xyzz*)o, "msm_acc_g1", "msm_reduce_g1"));
    TRY(msm_to_affine_canonical<Fq>(c, (const G1Xyzz*)o, (size_t)1, (G1Affine*)(o + sizeof(G2Xyzz))));
  } else {
    TRY(msm_run<Fq2>(c, tab ? b->table.as<G2Affine>() : b->pts.as<G2Affine>(), s, (G2Xyzz*)o, "msm_acc_g2", "msm_reduce_g2"));
    TRY(msm_to_affine_canonical<Fq2>(c, (const G2Xyzz*)o, (size_t)1, (G2Affine*)(o + sizeof(G2Xyzz))));
  }
  if (out) {
    CU(cudaMemcpyAsync(out, o + sizeof(G2Xyzz), b->group == 1 ? 64 : 128, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return 0;
}
static int msm_oneshot(zkfl_ctx* c, const uint8_t* bases, const uint8_t* scalars, size_t n, int group, uint8_t* out) {
  if (!scalars || !out) return fail(ZKFL_ERR_ARG, "bad argument");
  void* h = nullptr;
  TRY(msm_bases_upload(c, bases, n, group, &h));     // one MSM over these bases: no table (it would cost more than it saves)
  int rc = zkfl_msm_run(c, h, scalars, n, out);
  zkfl_msm_bases_free(h);
  return rc;
}
int zkfl_g1_msm(zkfl_ctx* c, const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t out[64]) { return msm_oneshot(c, bases, scalars, n, 1, out); }
int zkfl_g2_msm(zkfl_ctx* c, const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t out[128]) { return msm_oneshot(c, bases, scalars, n, 2, out); }

// ---- setup support
}  // extern "C"
extern "C" {
int zkfl_g1_mul_generator(zkfl_ctx* c, const uint8_t* scalars, size_t n, uint8_t* out) { return msm_gen_mul<Fq>(c, g1_generator(), scalars, n, out); }
int zkfl_g2_mul_generator(zkfl_ctx* c, const uint8_t* scalars, size_t n, uint8_t* out) { return msm_gen_mul<Fq2>(c, g2_generator(), scalars, n, out); }

// `zkey contribute`: every point of a section times one scalar
int zkfl_g1_scale_points(zkfl_ctx* c, const uint8_t* pts, size_t n, const uint8_t scalar[32], uint8_t* out) { return msm_point_scale<Fq>(c, pts, scalar, n, out); }
int zkfl_g2_scale_points(zkfl_ctx* c, const uint8_t* pts, size_t n, const uint8_t scalar[32], uint8_t* out) { return msm_point_scale<Fq2>(c, pts, scalar, n, out); }
// ---- single-proof forms of SURVEY 8b's list (one snarkjs / witness-calculator process each in the reference): B = 1 of the batch
int zkfl_wtns_calculate(zkfl_ctx* c, const zkfl_circuit* k, const zkfl_r1cs* r, const uint8_t* inputs, uint8_t* wtns_out, uint32_t* first_bad) {
  return zkfl_wtns_calculate_batch(c, k, r, inputs, 1, wtns_out, first_bad);
}
static const uint8_t* join_rs(const uint8_t* r, const uint8_t* s, uint8_t out[64]) {
  if (!r || !s) return nullptr;            // NULL -> CSPRNG, like snarkjs
  memcpy(out, r, 32); memcpy(out + 32, s, 32);
  return out;
}
int zkfl_groth16_prove(zkfl_ctx* c, const zkfl_zkey* z, const uint8_t* wtns, const uint8_t* r, const uint8_t* s, uint8_t proof_out[256],
                       uint8_t* public_out) {
  uint8_t rs[64];
  return zkfl_groth16_prove_batch(c, z, wtns, join_rs(r, s, rs), 1, proof_out, public_out);
}
int zkfl_groth16_full_prove(zkfl_ctx* c, const zkfl_circuit* k, const zkfl_zkey* z, const zkfl_r1cs* r1cs, const uint8_t* inputs,
                            const uint8_t* r, const uint8_t* s, uint8_t proof_out[256], uint8_t* public_out) {
  uint8_t rs[64];
  return zkfl_groth16_full_prove_batch(c, k, z, r1cs, inputs, join_rs(r, s, rs), 1, proof_out, public_out, nullptr);
}
// ---- snarkjs JSON shapes (proof.json / public.json) from the binary encodings; host code
static std::string dec_from_le32(const uint8_t* p) {
  uint32_t w[8]; memcpy(w, p, 32);
  std::string digits;
  for (;;) {
    uint64_t rem = 0; bool nz = false;
    for (int i = 7; i >= 0; i--) { uint64_t cur = (rem << 32) | w[i]; w[i] = (uint32_t)(cur / 1000000000u); rem = cur % 1000000000u; nz = nz || w[i]; }
    char buf[16];
    snprintf(buf, sizeof buf, nz ? "%09u" : "%u", (unsigned)rem);
    digits.insert(0, buf);
    if (!nz) break;
  }
  return digits;
}
static int put_json(const std::string& s, char* buf, size_t cap) {
  if (s.size() + 1 > cap) return fail(ZKFL_ERR_ARG, "buffer too small");
  memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}
int zkfl_proof_to_json(const uint8_t proof[256], char* buf, size_t cap) {
  if (!proof || !buf) return fail(ZKFL_ERR_ARG, "bad argument");
  std::string v[8];
  for (int i = 0; i < 8; i++) v[i] = "\"" + dec_from_le32(proof + 32 * i) + "\"";
  std::string j = "{\"pi_a\": [" + v[0] + ", " + v[1] + ", \"1\"], \"pi_b\": [[" + v[2] + ", " + v[3] + "], [" + v[4] + ", " + v[5] +
                  "], [\"1\", \"0\"]], \"pi_c\": [" + v[6] + ", " + v[7] + ", \"1\"], \"protocol\": \"groth16\", \"curve\": \"bn128\"}";
  return put_json(j, buf, cap);
}
int zkfl_public_to_json(const uint8_t* publics, uint32_t n_public, char* buf, size_t cap) {
  if ((!publics && n_public) || !buf) return fail(ZKFL_ERR_ARG, "bad argument");
  std::string j = "[";
  for (uint32_t i = 0; i < n_public; i++) j += (i ? ", \"" : "\"") + dec_from_le32(publics + 32 * (size_t)i) + "\"";
  j += "]";
  return put_json(j, buf, cap);
}
// makes `c`'s stream wait for everything queued so far on `other`'s stream (two contexts on one GPU working on
// half-batches concurrently: join before the end-of-step timestamp)
int zkfl_ctx_wait_other(zkfl_ctx* c, zkfl_ctx* other) {
  if (!c || !other || c->device != other->device) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  if (!other->ev_join) CU(cudaEventCreate(&other->ev_join));
  CU(cudaEventRecord(other->ev_join, other->stream));
  CU(cudaStreamWaitEvent(c->stream, other->ev_join, 0));
  return 0;
}
int zkfl_timer_begin(zkfl_ctx* c) {
  if (!c) return fail(ZKFL_ERR_ARG, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  if (!c->t0) { CU(cudaEventCreate(&c->t0)); CU(cudaEventCreate(&c->t1)); }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaEventRecord(c->t0, c->stream));
  return 0;
}
int zkfl_timer_end(zkfl_ctx* c, float* ms_out) {
  if (!c || !ms_out || !c->t0) return fail(ZKFL_ERR_ARG, "timer not started");
  CU(cudaEventRecord(c->t1, c->stream));
  CU(cudaEventSynchronize(c->t1));
  CU(cudaEventElapsedTime(ms_out, c->t0, c->t1));
  return 0;
}
static int bench_u32_kernel(zkfl_ctx* c, int which, size_t n_threads, uint32_t iters, float* ms_out);
int zkfl_bench_imad(zkfl_ctx* c, size_t n_threads, uint32_t iters, float* ms_out) { return bench_u32_kernel(c, 0, n_threads, iters, ms_out); }
int zkfl_bench_widemac(zkfl_ctx* c, size_t n_threads, uint32_t iters, float* ms_out) { return bench_u32_kernel(c, 1, n_threads, iters, ms_out); }
static int bench_u32_kernel(zkfl_ctx* c, int which, size_t n_threads, uint32_t iters, float* ms_out) {
  if (!c || !ms_out || n_threads == 0) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  DevBuf d;
  TRY(d.reserve(n_threads * 4));
  CU(cudaMemsetAsync(d.p, 0x5a, n_threads * 4, c->stream));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  if (which == 0) ZK_LAUNCH(k_bench_imad, n_threads, 256, c->stream, d.as<uint32_t>(), n_threads, 16u);
  else ZK_LAUNCH(k_bench_widemac, n_threads, 256, c->stream, d.as<uint32_t>(), n_threads, 16u);
  CU(cudaEventRecord(e0, c->stream));
  if (which == 0) ZK_LAUNCH(k_bench_imad, n_threads, 256, c->stream, d.as<uint32_t>(), n_threads, iters);
  else ZK_LAUNCH(k_bench_widemac, n_threads, 256, c->stream, d.as<uint32_t>(), n_threads, iters);
  CU(cudaEventRecord(e1, c->stream));
  CU(cudaEventSynchronize(e1));
  CU(cudaEventElapsedTime(ms_out, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}
int zkfl_bench_modmul(zkfl_ctx* c, size_t n_threads, uint32_t iters, float* ms_out) {
  if (!c || !ms_out || n_threads == 0) return fail(ZKFL_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  DevBuf d;
  TRY(d.reserve(n_threads * sizeof(Fq)));
  CU(cudaMemsetAsync(d.p, 0x5a, n_threads * sizeof(Fq), c->stream));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  ZK_LAUNCH(k_bench_modmul, n_threads, 256, c->stream, d.as<Fq>(), n_threads, 4u);  // warm-up
  CU(cudaEventRecord(e0, c->stream));
  ZK_LAUNCH(k_bench_modmul, n_threads, 256, c->stream, d.as<Fq>(), n_threads, iters);
  CU(cudaEventRecord(e1, c->stream));
  CU(cudaEventSynchronize(e1));
  CU(cudaEventElapsedTime(ms_out, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

}  // extern "C"
