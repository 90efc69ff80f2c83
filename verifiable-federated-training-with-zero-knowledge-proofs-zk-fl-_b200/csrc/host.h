// Host-side plumbing shared by the translation units of libzkfl.so: error reporting, device buffers, the context
// (one CUDA stream + grow-only workspace per context), stage timers, resident-artefact handles, and the entry points
// each .cu file offers to the others.  The library is split into several .cu files only so that nvcc can compile the
// kernel families in parallel (zkfl.cu: C ABI + artefact parsing + prove orchestration; witness.cu: witness / A.w,B.w /
// H polynomial; msm_g1.cu, msm_g2.cu: the Pippenger pipeline per group; verify.cu: batch verifier).
#pragma once
#include "types.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/zkfl.h"

using namespace zk;

int zk_fail(int code, const std::string& msg);     // sets zkfl_last_error() for this thread, returns code
uint64_t zk_launches_now();
static inline int fail(int code, const std::string& msg) { return zk_fail(code, msg); }

#define CU(expr)                                                                                     \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) return fail(ZKFL_ERR_CUDA, std::string(#expr) + ": " + zkrt::err_str(_e)); \
  } while (0)
#define TRY(expr)            \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    // grow with headroom so alternating batch sizes do not thrash
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { p = nullptr; return fail(ZKFL_ERR_NOMEM, "cudaMalloc(" + std::to_string(bytes) + ") failed"); }
    cap = bytes;
    return 0;
  }
  template <class T> T* as() const { return (T*)p; }
  ~DevBuf() { if (p) cudaFree(p); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

struct ProfRec { std::string name; cudaEvent_t e0, e1; uint64_t launches; };
struct ProfAgg { double ms = 0; uint64_t launches = 0; uint64_t calls = 0; };

struct zkfl_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool prof = false;
  std::vector<ProfRec> pending;
  std::map<std::string, ProfAgg> agg;
  std::vector<std::string> order;
  // workspace (grow-only)
  DevBuf w, abc, hsc, stage_in, stage_rs, aos;
  // counts / offsets of a bucket sort are still read by side-stream work (fix-up, reductions) while the main stream may
  // already run the NEXT sort: one set per sort of a proving pass (0: witness, 1: witness restricted to the B query, 2: H)
  DevBuf counts[3], offsets[3], cursors, chunk_sums, sorted[3], skey;   // sorted lists per sort too: the next sort runs on its own stream while the previous lists are accumulated
  DevBuf head[5], tail[5];   // chunk partials per MSM slot (read by that slot's reduction on its side stream)
  DevBuf fixq;               // large batches: ids of the buckets cut once / more than once (k_msm_fixup -> k_msm_fixup_apply)
  DevBuf heavy[5];            // heavy buckets of the current fix-up: [slots used | records (bucket, segment, segments, first slot)] + segment sums
  DevBuf v_ic, v_pub, v_proofs, v_t, v_g1, v_g2, v_flags, v_f, v_halves, v_ok;   // batch verifier
  DevBuf v_lines, v_rho, v_cps, v_sum[2], v_s, v_spart, v_tmul, v_sub, v_tree[2], v_misc;   // lane-cooperative / batched (RLC) verifier
  DevBuf aff_acc, aff_pre;   // batch-affine accumulation: running affine sums and prefix products, [slot group][lane]
  // five MSMs per proof batch (A, C, B1, H on G1; B2 on G2): own bucket / reduction buffers each, so the
  // latency-bound bucket reduction of one MSM runs on `side` while the next MSM accumulates on `stream`
  DevBuf buckets[5], Rs[5], Ts[5], lvl2[5], win[5];
  DevBuf red_main[5][2], red_pool[5][2];   // ping-pong buffers of the latency variant of the bucket reduction
  cudaStream_t side[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // one per MSM slot: the reductions are latency-bound and run concurrently
  cudaStream_t sort_stream = nullptr;      // the three sorts of a proving pass (L2-atomic bound) run beside the product-bound stages
  cudaEvent_t ev_in = nullptr, ev_hsc = nullptr, ev_sort[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_acc[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, ev_red[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  DevBuf res_g1, res_g2, t_g1, t_g2, pis, var, proofs, pubs, bad;
  // deferred checks of a proving call (constraint check of the HBM-resident witness, witness well-formedness): the kernels
  // run on the stream, their verdicts land in pinned host memory and are looked at after the call's final synchronisation,
  // so a check costs no mid-pipeline host round trip.  chk_host[0] = witness flags, chk_host[1 + b] = first violated row of b.
  uint32_t* chk_host = nullptr;
  size_t chk_cap = 0;
  uint32_t chk_B = 0;        // instances covered by the pending constraint check (0: none pending)
  bool chk_wtns = false;     // a witness well-formedness check is pending
  uint32_t w_wires = 0, w_B = 0;   // shape of the witness run_witness left in `w` (0: none)
  // per-key precomputations of the batch verifier, kept for the last four verification keys seen (a round alternates between
  // the three circuits' keys): byte-window tables of IC_0..IC_l and alpha (fixed-base scalar multiplications in 32 mixed
  // additions) and the Miller line tables of gamma, delta, beta
  struct VkCacheEntry { std::vector<uint8_t> key; DevBuf tabs, lines; uint64_t stamp = 0; };
  VkCacheEntry vk_cache[4];
  uint64_t vk_stamp = 0;
  int v_last_rlc = 0;        // 1 when the last zkfl_groth16_verify_batch was settled by the combined (random-linear-combination) check
  bool g2f_attr = false;     // k_msm_accumulate_chunks_g2f has been granted its dynamic shared memory on this device
  bool sort_attr = false;    // k_msm_sort_cta has been granted its dynamic shared memory on this device
  DevBuf msm_sc, msm_out, mask_w, mask_wb, mask_h, part_out, part_in;
  cudaEvent_t t0 = nullptr, t1 = nullptr, ev_join = nullptr;
};

struct Stage {
  zkfl_ctx* c; size_t idx = (size_t)-1; uint64_t l0; cudaStream_t st;
  Stage(zkfl_ctx* c_, const char* name, cudaStream_t stream = nullptr) : c(c_), st(stream ? stream : c_->stream) {
    if (!c->prof) return;
    ProfRec r; r.name = name; r.launches = 0;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    l0 = zk_launches_now();
    c->pending.push_back(r); idx = c->pending.size() - 1;
  }
  ~Stage() {
    if (idx == (size_t)-1) return;
    cudaEventRecord(c->pending[idx].e1, st);
    c->pending[idx].launches = zk_launches_now() - l0;
  }
};

struct Sec { const uint8_t* p; uint64_t len; };
int parse_sections(const uint8_t* d, size_t len, const char* magic, std::map<uint32_t, Sec>& out);

template <class T>
static inline int upload(zkfl_ctx* c, DevBuf& buf, const T* host, size_t count) {
  TRY(buf.reserve(count ? count * sizeof(T) : 16));
  if (count) CU(cudaMemcpyAsync(buf.p, host, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

struct CsrHost {
  std::vector<uint32_t> row_off, wire;
  std::vector<Fr> coef;
};
struct CsrBufs {
  DevBuf row_off, wire, coef;
  CsrDev dev() const { CsrDev d; d.row_off = row_off.as<uint32_t>(); d.wire = wire.as<uint32_t>(); d.coef = coef.as<Fr>(); return d; }
};
static inline int upload_csr(zkfl_ctx* c, const CsrHost& h, CsrBufs& b) {
  TRY(upload(c, b.row_off, h.row_off.data(), h.row_off.size()));
  TRY(upload(c, b.wire, h.wire.data(), h.wire.size()));
  TRY(upload(c, b.coef, h.coef.data(), h.coef.size()));
  return 0;
}

// ------------------------------------------------------------------------------------ handles
struct zkfl_circuit {
  zkfl_ctx* ctx;
  uint32_t n_wires, n_public, n_inputs, n_ops;
  DevBuf ops, lc_off, lc_wire, lc_coef, pos_in, pconst;
  std::vector<uint32_t> level_off;  // ops [level_off[k], level_off[k+1]) form dependency level k
  ProgramDev dev;
};
struct zkfl_r1cs {
  zkfl_ctx* ctx;
  uint32_t n_wires, n_constraints;
  CsrBufs A, B, C;
};
struct zkfl_zkey {
  zkfl_ctx* ctx;
  uint32_t n_vars, n_public, domain, log_n;
  CsrBufs A, B;
  DevBuf pA, pB1, pB2, pC, pH, skipB, tw_fwd, tw_inv, coset, tab_d1, tab_d2;
  uint32_t c_w = 0, c_h = 0;  // window sizes the precomputed tables were built for
  VkDev vk;
};
struct MsmBases {
  zkfl_ctx* ctx; int group; size_t n; DevBuf pts;
  // resident bases are constants of many MSMs: window-shifted table 2^(c*j) * P_i (index j*n + i) built once at load, so a
  // run needs ONE bucket set and one reduction instead of one per window (c_tab = 0: no table, per-window sets)
  DevBuf table; uint32_t c_tab = 0;
};

static inline uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* v = getenv(name);
  return v && *v ? (uint32_t)strtoul(v, nullptr, 10) : dflt;
}

// ---- msm_g1.cu (group-independent parts) and msm_g1.cu / msm_g2.cu (explicit instantiations for Fq / Fq2)
struct ReducePlan { uint32_t L1, L2, N1, N2; };
uint32_t accumulate_chunk(size_t entries);   // sorted entries per accumulation thread (by the size of the sort; ZKFL_MSM_CHUNK overrides)
MsmShape msm_shape(uint32_t m, uint32_t B, bool shared, uint32_t force_c = 0, uint32_t c_cap = 0);
ReducePlan reduce_plan(const MsmShape& s);
int msm_sort(zkfl_ctx* c, const Fr* scalars, const uint8_t* skip, const MsmShape& s, int gen = 0, cudaStream_t stream = nullptr);   // gen: which counts/offsets/list set; stream: default = the context's
int msm_sort_reserve(zkfl_ctx* c, const MsmShape& s, int gen);   // the allocations of msm_sort (cudaMalloc would serialise concurrent streams)
int msm_range_mask(zkfl_ctx* c, const uint8_t* base_skip, uint32_t m, uint32_t lo, uint32_t hi, uint8_t* out);
bool reduce_deep(const MsmShape& s);
int msm_reserve_reduce(zkfl_ctx* c, const MsmShape& s, int slot, size_t elem);
template <class F> int msm_accumulate(zkfl_ctx* c, const Affine<F>* bases, const MsmShape& s, int slot, const char* tag, int gen = 0, cudaStream_t stream = nullptr);
template <class F> int msm_reduce(zkfl_ctx* c, const MsmShape& s, int slot, Xyzz<F>* out, cudaStream_t stream, const char* tag, int gen = 0);
template <class F> int msm_run(zkfl_ctx* c, const Affine<F>* bases, const MsmShape& s, Xyzz<F>* out, const char* acc_tag, const char* red_tag);
template <class F> int msm_precompute_windows(zkfl_ctx* c, const Affine<F>* raw, uint32_t cnt, uint32_t cw, uint32_t W, Affine<F>* table);
template <class F> int msm_fixed_base_table(zkfl_ctx* c, const Affine<F>& base, Affine<F>* tab);
template <class F> int msm_to_affine_canonical(zkfl_ctx* c, const Xyzz<F>* in, size_t n, Affine<F>* out);
template <class F> int msm_sum_partials(zkfl_ctx* c, const Affine<F>* parts, uint32_t nparts, size_t part_stride, size_t n, Xyzz<F>* out);
template <class F> int msm_gen_mul(zkfl_ctx* c, const Affine<F>& gen, const uint8_t* scalars, size_t n, uint8_t* out);
template <class F> int msm_point_scale(zkfl_ctx* c, const uint8_t* pts, const uint8_t* scalar, size_t n, uint8_t* out);

// ---- witness.cu
int zk_aos_to_soa(zkfl_ctx* c, const Fr* src, Fr* dst, uint32_t n_elem, uint32_t B, uint32_t dst_elem_off);
int zk_soa_to_aos(zkfl_ctx* c, const Fr* src, Fr* dst, uint32_t n_elem, uint32_t B);
int zk_gather_wires(zkfl_ctx* c, const Fr* w, const uint32_t* wires, uint32_t n_sel, uint32_t B, Fr* out);
int run_witness(zkfl_ctx* c, const zkfl_circuit* circ, const uint8_t* inputs_host, uint32_t B);
int check_r1cs_device(zkfl_ctx* c, const zkfl_r1cs* r, uint32_t B, uint32_t* first_bad);
// deferred form: launch now (stream-ordered), judge after the caller's cudaStreamSynchronize
int check_r1cs_launch(zkfl_ctx* c, const zkfl_r1cs* r, uint32_t B);
int check_wtns_launch(zkfl_ctx* c, uint32_t n_wires, uint32_t B);   // every element < r and wire 0 == 1, on the [n_wires][B] witness in c->w
// ZKFL_ERR_ASSERT / ZKFL_ERR_ARG when a pending check failed; first_bad (may be NULL) has room for B entries
int checks_result(zkfl_ctx* c, uint32_t* first_bad, uint32_t B);
// a new pass starts: verdicts of an earlier asynchronous pass that nobody collected are dropped
static inline void checks_reset(zkfl_ctx* c) { c->chk_B = 0; c->chk_wtns = false; }
// witness in c->w -> H-MSM scalars in c->hsc; check != NULL: the constraint check is folded into the A.w / B.w pass (deferred verdict)
int run_h_poly(zkfl_ctx* c, const zkfl_zkey* z, uint32_t B, const zkfl_r1cs* check = nullptr);

// ---- field helpers (zkfl.cu)
bool fr_bytes_lt_mod(const uint8_t* p);
Fr fr_root_of_unity(int power);      // ffjavascript's 2^power-th root of unity (w[28] = 5^((r-1)/2^28), w[k] = w[k+1]^2), Montgomery

// the BN254 generators (1, 2) of G1 and the standard G2 generator, Montgomery affine
static inline G1Affine g1_generator() {
  G1Affine g; g.x = Fq::zero(); g.x.v[0] = 1; g.x = g.x.to_mont(); g.y = Fq::zero(); g.y.v[0] = 2; g.y = g.y.to_mont();
  return g;
}
static inline G2Affine g2_generator() {
  static const uint32_t W[4][8] = {{0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu},
                                   {0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u},
                                   {0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u},
                                   {0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u}};
  Fq v[4];
  for (int k = 0; k < 4; k++) { for (int i = 0; i < 8; i++) v[k].v[i] = W[k][i]; v[k] = v[k].to_mont(); }
  G2Affine g; g.x.a = v[0]; g.x.b = v[1]; g.y.a = v[2]; g.y.b = v[3];
  return g;
}
