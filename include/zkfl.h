/* zkfl.h -- C ABI of libzkfl.so, the B200 (sm_100a) Groth16/BN254 proving backend.
 *
 * Drop-in boundary: the reference (/root/reference) has no FFI of its own; its tests reach the
 * prover by spawning the snarkjs / circom tool-chain (`runCommand`, tests/full_system_simulation.mjs:108-115).
 * Each entry point below names the tool invocation (reference call site) it replaces.  A Node N-API
 * addon, the Python host package and the CLI shim all bind exactly these symbols.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * code on failure (message: zkfl_last_error()); no exception crosses the ABI; the caller owns all
 * buffers; handles are opaque and released by the matching *_free.  A zkfl_ctx is bound to one CUDA
 * device and is NOT thread-safe (one ctx per GPU per host thread); zkfl_circuit / zkfl_zkey handles are read-only
 * after loading and may be used by several contexts of the same device concurrently (one resident copy of the key
 * tables).  There is no CPU fallback: without a usable CUDA device zkfl_ctx_create fails.
 *
 * Encodings: field elements are 32-byte little-endian canonical integers (the `.wtns` encoding);
 * a proof is 256 bytes: pi_a (x,y) | pi_b (x.c0,x.c1,y.c0,y.c1) | pi_c (x,y), affine canonical
 * little-endian -- the numbers snarkjs prints in proof.json.  `.zkey` / `.r1cs` are the iden3 binary
 * formats snarkjs 0.7 reads and writes; `.zkwp` is this library's compiled witness program (the
 * replacement for the circom-generated `<name>_js/<name>.wasm`).
 */
#ifndef ZKFL_H
#define ZKFL_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zkfl_ctx zkfl_ctx;
typedef struct zkfl_circuit zkfl_circuit; /* compiled witness program (.zkwp)  ~ circom .wasm */
typedef struct zkfl_zkey zkfl_zkey;       /* proving key resident in HBM       ~ snarkjs .zkey */
typedef struct zkfl_r1cs zkfl_r1cs;       /* constraint system resident in HBM ~ circom .r1cs */

#define ZKFL_OK 0
#define ZKFL_ERR_ARG (-1)
#define ZKFL_ERR_FORMAT (-2)
#define ZKFL_ERR_CUDA (-3)
#define ZKFL_ERR_NOMEM (-4)
#define ZKFL_ERR_ASSERT (-5) /* a circuit `===` failed: circom's "Assert Failed" */

const char* zkfl_last_error(void);
const char* zkfl_version(void);

/* ---- context ------------------------------------------------------------------------------- */
int zkfl_ctx_create(int device, zkfl_ctx** out);
void zkfl_ctx_free(zkfl_ctx* ctx);

/* ---- artefacts ----------------------------------------------------------------------------- */
/* replaces loading `<name>.wasm` in generate_witness.cjs (tests/full_system_simulation.mjs:760-762) */
int zkfl_circuit_load(zkfl_ctx* ctx, const uint8_t* zkwp, size_t len, zkfl_circuit** out);
void zkfl_circuit_free(zkfl_circuit* c);
/* info[0..3] = n_wires, n_public, n_inputs, n_ops */
int zkfl_circuit_info(const zkfl_circuit* c, uint32_t info[4]);

/* replaces snarkjs reading `<name>_final.zkey` in `groth16 prove` (tests/full_system_simulation.mjs:773-775);
 * parses the snarkjs layout and uploads coefficients and bases once */
int zkfl_zkey_load(zkfl_ctx* ctx, const uint8_t* zkey, size_t len, zkfl_zkey** out);
/* the same key for ONE proof split over nparts GPUs (zkfl_groth16_msm_partials, BASELINE.json configs[4]; the reference has no
 * multi-GPU form: this is the north-star's "single large proofs split ... across GPUs"): window tables sized for a rank's share
 * of the points, so that the per-rank bucket reduction shrinks with the number of ranks.  Proof bytes do not depend on it. */
int zkfl_zkey_load_split(zkfl_ctx* ctx, const uint8_t* zkey, size_t len, uint32_t nparts, zkfl_zkey** out);
void zkfl_zkey_free(zkfl_zkey* z);
/* info[0..2] = n_vars, n_public, domain_size */
int zkfl_zkey_info(const zkfl_zkey* z, uint32_t info[3]);

/* `.r1cs` for the constraint check the circom witness calculator performs at every `===` */
int zkfl_r1cs_load(zkfl_ctx* ctx, const uint8_t* r1cs, size_t len, zkfl_r1cs** out);
void zkfl_r1cs_free(zkfl_r1cs* r);

/* ---- witness: `node generate_witness.cjs wasm input.json out.wtns`, `snarkjs wtns calculate`
 *      (tests/full_system_simulation.mjs:760-762, tests/test_secureagg.cjs:108-118), batched ------- */
/* inputs: B x n_inputs field elements in the circuit's input-declaration order (the flattened
 * input.json); wtns_out: B x n_wires field elements (section 2 of B `.wtns` files).
 * If r1cs != NULL every constraint is checked on the GPU and first_bad[b] (B entries, may be NULL)
 * receives the first violated constraint index or 0xFFFFFFFF; any violation -> ZKFL_ERR_ASSERT. */
int zkfl_wtns_calculate_batch(zkfl_ctx* ctx, const zkfl_circuit* c, const zkfl_r1cs* r1cs, const uint8_t* inputs,
                              int B, uint8_t* wtns_out, uint32_t* first_bad);
int zkfl_r1cs_check_batch(zkfl_ctx* ctx, const zkfl_r1cs* r1cs, const uint8_t* wtns, int B, uint32_t* first_bad);
/* Runs a witness program for B instances and returns only `n_sel` selected wires (out: B x n_sel field elements).
 * This is how the OFF-circuit commitment pipeline of the reference (vectorHash / buildMerkleTree / derivePairwiseMask,
 * tests/full_system_simulation.mjs:139-238, computed there with circomlibjs on the CPU) runs on the GPU: the commitments
 * of all clients are the wires of a "commitment program" compiled by the same front-end (zkfl_b200/commitments.py). */
int zkfl_wtns_eval_wires(zkfl_ctx* ctx, const zkfl_circuit* c, const uint8_t* inputs, int B, const uint32_t* wires,
                         uint32_t n_sel, uint8_t* out);

/* ---- prove: `snarkjs groth16 prove zkey wtns proof.json public.json`
 *      (tests/full_system_simulation.mjs:773-775 and five more call sites), batched ---- */
/* wtns: B x n_vars; rs: B x 64 bytes (r then s, canonical, < r) or NULL for OS randomness
 * (snarkjs draws r,s from a CSPRNG); proofs_out: B x 256; publics_out: B x n_public x 32 (may be NULL).
 * The witness must be well-formed (every element < r, wire 0 == 1; checked on the device) else ZKFL_ERR_ARG. */
int zkfl_groth16_prove_batch(zkfl_ctx* ctx, const zkfl_zkey* z, const uint8_t* wtns, const uint8_t* rs, int B,
                             uint8_t* proofs_out, uint8_t* publics_out);
/* `snarkjs.groth16.fullProve(input, wasm, zkey)` batched: witness stays in HBM between the two steps.
 * Inputs must be reduced mod r (else ZKFL_ERR_ARG).  r1cs != NULL: every `===` is checked on the HBM-resident witness in
 * the same pass (circom aborts fullProve on a failed assert): any violation -> ZKFL_ERR_ASSERT, proofs_out zeroed,
 * first_bad[b] (B entries, may be NULL) = first violated constraint or 0xFFFFFFFF.  r1cs == NULL: no check (the caller
 * vouches for the inputs, e.g. a benchmark replaying known-good instances). */
int zkfl_groth16_full_prove_batch(zkfl_ctx* ctx, const zkfl_circuit* c, const zkfl_zkey* z, const zkfl_r1cs* r1cs,
                                  const uint8_t* inputs, const uint8_t* rs, int B, uint8_t* proofs_out, uint8_t* publics_out,
                                  uint32_t* first_bad);
/* device-resident variant for steady-state measurement: stage() copies inputs and rs to HBM once,
 * run() proves from HBM leaving proofs in HBM, fetch() copies B x 256 proof bytes back. */
int zkfl_full_prove_stage(zkfl_ctx* ctx, const zkfl_circuit* c, const zkfl_zkey* z, const uint8_t* inputs,
                          const uint8_t* rs, int B);
int zkfl_full_prove_run(zkfl_ctx* ctx, const zkfl_circuit* c, const zkfl_zkey* z, const zkfl_r1cs* r1cs /* may be NULL */, int B);
/* synchronises; reports the constraint check of the run (ZKFL_ERR_ASSERT, first_bad as above; first_bad may be NULL) */
int zkfl_full_prove_fetch(zkfl_ctx* ctx, int B, uint8_t* proofs_out, uint32_t* first_bad);

/* ---- one large proof split across GPUs (BASELINE.json configs[4]): rank `part` of `nparts` runs the five
 *      multi-scalar multiplications over ITS point range [part*m/nparts, (part+1)*m/nparts) and returns the partial
 *      sums; the ranks all-gather the partials (NCCL / gloo: 384 B per proof per rank) and any rank finishes with
 *      zkfl_groth16_finalize, which adds the partials (the group law is no NCCL reduction op) and applies the blinding.
 *      partials layout per call: [A(64) x B | B1 x B | C x B | H x B | B2(128) x B], affine canonical.
 *      wtns == NULL: the witness is the one the last witness calculation on this context left in HBM (zkfl_wtns_calculate_batch
 *      with wtns_out == NULL keeps it there) -- nothing is uploaded.  partials_out / partials may be HOST or DEVICE buffers: with
 *      device buffers the all-gather runs device-to-device (NCCL over NVLink) and nothing is staged through the host. */
int zkfl_groth16_msm_partials(zkfl_ctx* ctx, const zkfl_zkey* z, const uint8_t* wtns, int B, uint32_t part, uint32_t nparts,
                              uint8_t* partials_out /* B x 384 */);
int zkfl_groth16_finalize(zkfl_ctx* ctx, const zkfl_zkey* z, const uint8_t* partials /* nparts x B x 384 */, uint32_t nparts,
                          const uint8_t* rs, int B, uint8_t* proofs_out /* B x 256 */);

/* ---- single-proof forms (what one `generate_witness.cjs` / `snarkjs groth16 prove` process does,
 *      tests/full_system_simulation.mjs:760-762,773-775): the batch entry points above with B = 1.
 *      r, s: 32-byte canonical blinding scalars, or NULL for random ones (what snarkjs does). ------------------------- */
int zkfl_wtns_calculate(zkfl_ctx* ctx, const zkfl_circuit* c, const zkfl_r1cs* r1cs, const uint8_t* inputs, uint8_t* wtns_out,
                        uint32_t* first_bad);
int zkfl_groth16_prove(zkfl_ctx* ctx, const zkfl_zkey* z, const uint8_t* wtns, const uint8_t* r, const uint8_t* s,
                       uint8_t proof_out[256], uint8_t* public_out);
int zkfl_groth16_full_prove(zkfl_ctx* ctx, const zkfl_circuit* c, const zkfl_zkey* z, const zkfl_r1cs* r1cs /* may be NULL */,
                            const uint8_t* inputs, const uint8_t* r, const uint8_t* s, uint8_t proof_out[256], uint8_t* public_out);
/* snarkjs's proof.json / public.json texts from the binary encodings (decimal strings, "protocol": "groth16", "curve": "bn128") */
int zkfl_proof_to_json(const uint8_t proof[256], char* buf, size_t cap);
int zkfl_public_to_json(const uint8_t* publics, uint32_t n_public, char* buf, size_t cap);

/* ---- verify: `snarkjs groth16 verify vkey.json public.json proof.json`
 *      (tests/full_system_simulation.mjs:865-868,975-978,1116-1119). Host-side pairing check (SURVEY 2.4 row V1:
 *      verification stays on the CPU). All points affine canonical little-endian as in vkey.json / proof.json:
 *      alpha1 64 B, beta2/gamma2/delta2 128 B (x.c0,x.c1,y.c0,y.c1), ic (n_public+1) x 64 B, publics n_public x 32 B.
 *      *ok = 1 iff e(A,B) = e(alpha,beta) e(vk_x,gamma) e(C,delta) and every public signal is < r. ---------------- */
int zkfl_groth16_verify(const uint8_t* alpha1, const uint8_t* beta2, const uint8_t* gamma2, const uint8_t* delta2,
                        const uint8_t* ic, const uint8_t* publics, uint32_t n_public, const uint8_t* proof, int* ok);

/* ---- batch verify on the GPU (Server.verifyBalanceProof / verifyTrainingProof /
 *      verifySecureAggregationProof, tests/full_system_simulation.mjs:848-1131 -- one `snarkjs groth16 verify` child process
 *      per proof there).  B proofs under ONE verification key, same encodings as zkfl_groth16_verify;
 *      publics: B x n_public x 32 B, proofs: B x 256 B, ok: B x int32 (1 = valid, 0 = invalid or malformed proof).
 *      Per proof: vk_x, three Miller loops, and the final exponentiation, spread over threads (csrc/pairing.cuh). ------- */
int zkfl_groth16_verify_batch(zkfl_ctx* ctx, const uint8_t* alpha1, const uint8_t* beta2, const uint8_t* gamma2,
                              const uint8_t* delta2, const uint8_t* ic, uint32_t n_public, const uint8_t* publics,
                              const uint8_t* proofs, int B, int32_t* ok);

/* dev / test hook: intermediate buffers of the last zkfl_groth16_verify_batch ("v_t", "v_g1", "v_g2", "v_flags", "v_f", "v_halves") */
int zkfl_debug_read(zkfl_ctx* ctx, const char* name, void* out, size_t bytes);
/* host-side consistency check of the verifier's two Fq12 views (products, inverse, Frobenius maps); 0 = ok, else the failing check */
int zkfl_debug_pairing_selftest(void);

/* ---- standalone multi-scalar multiplication (BASELINE.json: "G1 MSM pts/s at 2^20") ----------- */
/* bases: n affine points, Montgomery little-endian (zkey point layout, 64 B G1 / 128 B G2);
 * scalars: n x 32 B canonical; out: affine canonical (64 / 128 B). */
int zkfl_g1_msm(zkfl_ctx* ctx, const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t out[64]);
int zkfl_g2_msm(zkfl_ctx* ctx, const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t out[128]);
/* resident variant: bases uploaded once (they are per-circuit constants), then repeated MSMs */
int zkfl_msm_bases_load(zkfl_ctx* ctx, const uint8_t* bases, size_t n, int group /*1|2*/, void** handle);
void zkfl_msm_bases_free(void* handle);
int zkfl_msm_run(zkfl_ctx* ctx, void* handle, const uint8_t* scalars /* host, or NULL = reuse staged */, size_t n,
                 uint8_t* out);

/* ---- setup support: `snarkjs groth16 setup` (tests/full_system_simulation.mjs:714-717) needs
 *      k_i * G for every key scalar; out = affine Montgomery (zkey point layout) ------------------ */
int zkfl_g1_mul_generator(zkfl_ctx* ctx, const uint8_t* scalars, size_t n, uint8_t* out /* n x 64 */);
int zkfl_g2_mul_generator(zkfl_ctx* ctx, const uint8_t* scalars, size_t n, uint8_t* out /* n x 128 */);

/* `snarkjs groth16 setup <r1cs> <ptau> <zkey>` (tests/full_system_simulation.mjs:714-717) on the GPU: Lagrange basis at tau, column
 * sums of A / B / C, key scalars and all scalar multiplications run on the device; the result is the complete `.zkey` (snarkjs
 * section layout).  The R1CS comes as coordinate lists per matrix k = 0 (A), 1 (B), 2 (C): rows[k][i], wires[k][i] and
 * cidx[k][i] (index into coef_table: n_coef distinct coefficient values, 32 B canonical) for i < nnz[k].
 * toxic = tau | alpha | beta | delta (4 x 32 B canonical, non-zero): WHOEVER KNOWS THEM CAN FORGE PROOFS -- callers draw them from a
 * CSPRNG and forget them (zkfl_b200.zkey_setup), fixed values are for tests.  Call with zkey_out == NULL to learn *zkey_len. */
int zkfl_groth16_setup(zkfl_ctx* ctx, uint32_t n_wires, uint32_t n_public, uint32_t n_constraints, const uint32_t* const rows[3],
                       const uint32_t* const wires[3], const uint32_t* const cidx[3], const size_t nnz[3], const uint8_t* coef_table,
                       uint32_t n_coef, const uint8_t toxic[128], uint8_t* zkey_out, size_t cap, size_t* zkey_len);
/* `snarkjs zkey contribute <in> <out> --name= -e=` (tests/full_system_simulation.mjs:726-731) multiplies delta by a fresh
 * secret d and the C / H sections by 1/d: out[i] = scalar * pts[i], points affine Montgomery (zkey layout), scalar canonical */
int zkfl_g1_scale_points(zkfl_ctx* ctx, const uint8_t* pts, size_t n, const uint8_t scalar[32], uint8_t* out /* n x 64 */);
int zkfl_g2_scale_points(zkfl_ctx* ctx, const uint8_t* pts, size_t n, const uint8_t scalar[32], uint8_t* out /* n x 128 */);

/* ---- Server.aggregateUpdates (tests/full_system_simulation.mjs:1137-1199) on the device: per model coordinate the field sum of
 *      the accepted clients' masked updates (the pairwise masks cancel), the signed decode (sums above r/2 are negative integers,
 *      converted like JavaScript's Number(BigInt)), the mean over the accepted clients and the SGD step
 *      model_out[j] = model_in[j] - learning_rate * mean[j] (IEEE doubles, product and difference rounded separately as in JS).
 *      masked: n_clients x dim x 32 B canonical (host or device); accept: n_clients bytes (NULL = all); no accepted client ->
 *      ZKFL_ERR_ARG ("No verified updates to aggregate!": the reference returns null). ------------------------------------ */
int zkfl_aggregate_updates(zkfl_ctx* ctx, const uint8_t* masked, const uint8_t* accept, uint32_t n_clients, uint32_t dim,
                           double learning_rate, const double* model_in, uint8_t* agg_field_out /* dim x 32, may be NULL */,
                           double* agg_mean_out /* dim */, double* model_out /* dim */, uint32_t* n_accepted /* may be NULL */);

/* ---- measurement --------------------------------------------------------------------------- */
/* number of CUDA kernels this library has launched in this process */
uint64_t zkfl_launch_count(void);
/* per-stage device timing (CUDA events on the context's stream). enable, run, then read a text
 * table "stage ms launches\n..." into buf. */
int zkfl_prof_enable(zkfl_ctx* ctx, int on);
int zkfl_prof_read(zkfl_ctx* ctx, char* buf, size_t cap);
/* two contexts on the same device (e.g. two half-batches in flight): make ctx's stream wait for other's queued work */
int zkfl_ctx_wait_other(zkfl_ctx* ctx, zkfl_ctx* other);
/* device timer on the context's stream: begin() synchronises and records, end() records, waits, returns ms */
int zkfl_timer_begin(zkfl_ctx* ctx);
int zkfl_timer_end(zkfl_ctx* ctx, float* ms_out);
/* integer-pipe microbenchmarks (roofline denominators), each returns the kernel's ms:
 *   modmul: n threads x iters x 2 dependent Montgomery products (136 MAC each)
 *   imad:   n threads x iters x 8 independent 32-bit multiply-adds */
int zkfl_bench_modmul(zkfl_ctx* ctx, size_t n_threads, uint32_t iters, float* ms_out);
int zkfl_bench_imad(zkfl_ctx* ctx, size_t n_threads, uint32_t iters, float* ms_out);
/*   widemac: n threads x iters x 4 fused 32x32->64 multiply-accumulates (IMAD.WIDE.U32.X), the unit "MAC" of the rooflines */
int zkfl_bench_widemac(zkfl_ctx* ctx, size_t n_threads, uint32_t iters, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* ZKFL_H */
