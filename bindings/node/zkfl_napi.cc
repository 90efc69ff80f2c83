// zkfl_napi.cc -- Node.js addon (plain N-API, no node-addon-api dependency) over the C ABI of libzkfl.so (include/zkfl.h).
//
// What it replaces in the reference: the three child processes per proof of tests/full_system_simulation.mjs
// (`node generate_witness.cjs` :760-762, `npx snarkjs groth16 prove` :773-775, `npx snarkjs groth16 verify` :865-868) and the
// north star's `snarkjs.groth16.fullProve(input, wasm, zkey)`.  zkfl_snarkjs.mjs (next to this file) puts the snarkjs names
// and result shapes on top of these functions.
//
// Exposed to JavaScript (all synchronous: one call = one GPU pass; a phase of the federated round proves ALL its clients in one call):
//   loadCircuit(zkwp: Buffer, r1cs?: Buffer)           -> External   zkfl_circuit_load (+ zkfl_r1cs_load for the `===` check)
//   loadZkey(zkey: Buffer)                             -> External   zkfl_zkey_load
//   circuitInfo(circuit) / zkeyInfo(zkey)              -> {nWires, nPublic, nInputs} / {nVars, nPublic, domain}
//   wtnsCalculateBatch(circuit, inputs: Buffer, B)     -> Buffer(B * nWires * 32)        zkfl_wtns_calculate_batch
//   proveBatch(zkey, wtns: Buffer, B, rs?: Buffer)     -> {proofs, publics}              zkfl_groth16_prove_batch
//   fullProveBatch(circuit, zkey, inputs: Buffer, B, rs?: Buffer) -> {proofs: Buffer(B*256), publics: Buffer(B*nPublic*32)}
//                                                                                         zkfl_groth16_full_prove_batch
//   verify(vk: {alpha1,beta2,gamma2,delta2,ic: Buffer, nPublic}, publics: Buffer, proof: Buffer) -> boolean   zkfl_groth16_verify
//   verifyBatch(vk, publics: Buffer, proofs: Buffer, B) -> boolean[]                     zkfl_groth16_verify_batch
//   proofToJson(proof: Buffer) / publicToJson(publics: Buffer, nPublic) -> string        zkfl_proof_to_json / zkfl_public_to_json
// Errors: a JavaScript Error carrying zkfl_last_error(); a failed circuit `===` has code "ZKFL_ASSERT" (circom's "Assert Failed").
// There is no CPU fallback: loading the addon on a machine without a CUDA device throws from the first call.
//
// Build (needs Node headers, absent from the image this repository is developed on, so this file is compiled only against
// the declaration stub tests/stubs/node_api.h by the CPU test-suite):  see binding.gyp.
#include <node_api.h>

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

extern "C" {
#include "zkfl.h"
}

namespace {

zkfl_ctx* g_ctx = nullptr;   // one context per process = one per GPU (ZKFL_DEVICE / LOCAL_RANK select the device)

struct CircuitHandle { zkfl_circuit* c; zkfl_r1cs* r; };

napi_value throw_zkfl(napi_env env, int rc) {
  napi_throw_error(env, rc == ZKFL_ERR_ASSERT ? "ZKFL_ASSERT" : "ZKFL_ERROR", zkfl_last_error());
  return nullptr;
}
napi_value throw_type(napi_env env, const char* msg) {
  napi_throw_type_error(env, nullptr, msg);
  return nullptr;
}
bool ensure_ctx(napi_env env) {
  if (g_ctx) return true;
  const char* d = getenv("ZKFL_DEVICE");
  if (!d) d = getenv("LOCAL_RANK");
  int rc = zkfl_ctx_create(d ? atoi(d) : 0, &g_ctx);
  if (rc) { throw_zkfl(env, rc); return false; }
  return true;
}
// argument helpers: false after having thrown
bool get_args(napi_env env, napi_callback_info info, size_t want_min, size_t* argc, napi_value* argv) {
  if (napi_get_cb_info(env, info, argc, argv, nullptr, nullptr) != napi_ok) return false;
  if (*argc < want_min) { throw_type(env, "too few arguments"); return false; }
  return true;
}
bool get_buffer(napi_env env, napi_value v, const uint8_t** data, size_t* len) {
  bool is = false;
  if (napi_is_buffer(env, v, &is) != napi_ok || !is) { throw_type(env, "Buffer expected"); return false; }
  void* p = nullptr;
  if (napi_get_buffer_info(env, v, &p, len) != napi_ok) return false;
  *data = (const uint8_t*)p;
  return true;
}
bool get_optional_buffer(napi_env env, size_t argc, napi_value* argv, size_t idx, const uint8_t** data, size_t* len) {
  *data = nullptr; *len = 0;
  if (idx >= argc) return true;
  napi_valuetype t;
  if (napi_typeof(env, argv[idx], &t) != napi_ok) return false;
  if (t == napi_undefined || t == napi_null) return true;
  return get_buffer(env, argv[idx], data, len);
}
bool get_int(napi_env env, napi_value v, int32_t* out) {
  if (napi_get_value_int32(env, v, out) != napi_ok) { throw_type(env, "integer expected"); return false; }
  return true;
}
template <class T> bool get_external(napi_env env, napi_value v, T** out) {
  void* p = nullptr;
  if (napi_get_value_external(env, v, &p) != napi_ok || !p) { throw_type(env, "handle expected"); return false; }
  *out = (T*)p;
  return true;
}
napi_value new_buffer(napi_env env, size_t len, uint8_t** data) {
  napi_value b; void* p = nullptr;
  if (napi_create_buffer(env, len, &p, &b) != napi_ok) return nullptr;
  *data = (uint8_t*)p;
  return b;
}
void set_u32(napi_env env, napi_value obj, const char* key, uint32_t v) {
  napi_value n;
  napi_create_uint32(env, v, &n);
  napi_set_named_property(env, obj, key, n);
}
bool named_buffer(napi_env env, napi_value obj, const char* key, const uint8_t** data, size_t* len) {
  napi_value v;
  if (napi_get_named_property(env, obj, key, &v) != napi_ok) { throw_type(env, "verification key field missing"); return false; }
  return get_buffer(env, v, data, len);
}

void free_circuit(napi_env, void* data, void*) {
  CircuitHandle* h = (CircuitHandle*)data;
  if (h->r) zkfl_r1cs_free(h->r);
  if (h->c) zkfl_circuit_free(h->c);
  delete h;
}
void free_zkey(napi_env, void* data, void*) { zkfl_zkey_free((zkfl_zkey*)data); }

// loadCircuit(zkwp, r1cs?)
napi_value LoadCircuit(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2];
  if (!get_args(env, info, 1, &argc, argv) || !ensure_ctx(env)) return nullptr;
  const uint8_t *prog, *r1 = nullptr; size_t plen, rlen = 0;
  if (!get_buffer(env, argv[0], &prog, &plen) || !get_optional_buffer(env, argc, argv, 1, &r1, &rlen)) return nullptr;
  CircuitHandle* h = new CircuitHandle{nullptr, nullptr};
  int rc = zkfl_circuit_load(g_ctx, prog, plen, &h->c);
  if (!rc && r1) rc = zkfl_r1cs_load(g_ctx, r1, rlen, &h->r);
  if (rc) { free_circuit(env, h, nullptr); return throw_zkfl(env, rc); }
  napi_value ext;
  if (napi_create_external(env, h, free_circuit, nullptr, &ext) != napi_ok) { free_circuit(env, h, nullptr); return nullptr; }
  return ext;
}
// loadZkey(zkey)
napi_value LoadZkey(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1];
  if (!get_args(env, info, 1, &argc, argv) || !ensure_ctx(env)) return nullptr;
  const uint8_t* d; size_t len;
  if (!get_buffer(env, argv[0], &d, &len)) return nullptr;
  zkfl_zkey* z = nullptr;
  int rc = zkfl_zkey_load(g_ctx, d, len, &z);
  if (rc) return throw_zkfl(env, rc);
  napi_value ext;
  if (napi_create_external(env, z, free_zkey, nullptr, &ext) != napi_ok) { zkfl_zkey_free(z); return nullptr; }
  return ext;
}
napi_value CircuitInfo(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1];
  CircuitHandle* h;
  if (!get_args(env, info, 1, &argc, argv) || !get_external(env, argv[0], &h)) return nullptr;
  uint32_t v[4];
  int rc = zkfl_circuit_info(h->c, v);
  if (rc) return throw_zkfl(env, rc);
  napi_value o; napi_create_object(env, &o);
  set_u32(env, o, "nWires", v[0]); set_u32(env, o, "nPublic", v[1]); set_u32(env, o, "nInputs", v[2]);
  return o;
}
napi_value ZkeyInfo(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1];
  zkfl_zkey* z;
  if (!get_args(env, info, 1, &argc, argv) || !get_external(env, argv[0], &z)) return nullptr;
  uint32_t v[3];
  int rc = zkfl_zkey_info(z, v);
  if (rc) return throw_zkfl(env, rc);
  napi_value o; napi_create_object(env, &o);
  set_u32(env, o, "nVars", v[0]); set_u32(env, o, "nPublic", v[1]); set_u32(env, o, "domain", v[2]);
  return o;
}
// wtnsCalculateBatch(circuit, inputs, B) -> Buffer; throws ZKFL_ASSERT when a `===` fails (the r1cs given to loadCircuit is checked)
napi_value WtnsCalculateBatch(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  CircuitHandle* h; const uint8_t* in; size_t in_len; int32_t B;
  if (!get_args(env, info, 3, &argc, argv) || !get_external(env, argv[0], &h) || !get_buffer(env, argv[1], &in, &in_len) ||
      !get_int(env, argv[2], &B))
    return nullptr;
  uint32_t ci[4];
  if (zkfl_circuit_info(h->c, ci)) return throw_zkfl(env, ZKFL_ERR_ARG);
  if (B <= 0 || in_len != (size_t)B * ci[2] * 32) return throw_type(env, "inputs must hold B * nInputs field elements of 32 bytes");
  uint8_t* out;
  napi_value buf = new_buffer(env, (size_t)B * ci[0] * 32, &out);
  if (!buf) return nullptr;
  int rc = zkfl_wtns_calculate_batch(g_ctx, h->c, h->r, in, B, out, nullptr);
  if (rc) return throw_zkfl(env, rc);
  return buf;
}
napi_value make_result(napi_env env, napi_value proofs, napi_value pubs) {
  napi_value o; napi_create_object(env, &o);
  napi_set_named_property(env, o, "proofs", proofs);
  napi_set_named_property(env, o, "publics", pubs);
  return o;
}
// proveBatch(zkey, wtns, B, rs?)
napi_value ProveBatch(napi_env env, napi_callback_info info) {
  size_t argc = 4; napi_value argv[4];
  zkfl_zkey* z; const uint8_t *w, *rs; size_t wlen, rslen; int32_t B;
  if (!get_args(env, info, 3, &argc, argv) || !get_external(env, argv[0], &z) || !get_buffer(env, argv[1], &w, &wlen) ||
      !get_int(env, argv[2], &B) || !get_optional_buffer(env, argc, argv, 3, &rs, &rslen))
    return nullptr;
  uint32_t zi[3];
  if (zkfl_zkey_info(z, zi)) return throw_zkfl(env, ZKFL_ERR_ARG);
  if (B <= 0 || wlen != (size_t)B * zi[0] * 32) return throw_type(env, "Invalid witness length");
  if (rs && rslen != (size_t)B * 64) return throw_type(env, "rs must hold B * 64 bytes");
  uint8_t *p, *q;
  napi_value proofs = new_buffer(env, (size_t)B * 256, &p), pubs = new_buffer(env, (size_t)B * zi[1] * 32, &q);
  if (!proofs || !pubs) return nullptr;
  int rc = zkfl_groth16_prove_batch(g_ctx, z, w, rs, B, p, q);
  if (rc) return throw_zkfl(env, rc);
  return make_result(env, proofs, pubs);
}
// fullProveBatch(circuit, zkey, inputs, B, rs?): witness + constraint check + prove in one GPU pass
napi_value FullProveBatch(napi_env env, napi_callback_info info) {
  size_t argc = 5; napi_value argv[5];
  CircuitHandle* h; zkfl_zkey* z; const uint8_t *in, *rs; size_t in_len, rslen; int32_t B;
  if (!get_args(env, info, 4, &argc, argv) || !get_external(env, argv[0], &h) || !get_external(env, argv[1], &z) ||
      !get_buffer(env, argv[2], &in, &in_len) || !get_int(env, argv[3], &B) || !get_optional_buffer(env, argc, argv, 4, &rs, &rslen))
    return nullptr;
  uint32_t ci[4], zi[3];
  if (zkfl_circuit_info(h->c, ci) || zkfl_zkey_info(z, zi)) return throw_zkfl(env, ZKFL_ERR_ARG);
  if (B <= 0 || in_len != (size_t)B * ci[2] * 32) return throw_type(env, "inputs must hold B * nInputs field elements of 32 bytes");
  if (rs && rslen != (size_t)B * 64) return throw_type(env, "rs must hold B * 64 bytes");
  if (!h->r) return throw_type(env, "fullProve needs the circuit's .r1cs (pass it to loadCircuit): the `===` check is never skipped silently");
  uint8_t *p, *q;
  napi_value proofs = new_buffer(env, (size_t)B * 256, &p), pubs = new_buffer(env, (size_t)B * zi[1] * 32, &q);
  if (!proofs || !pubs) return nullptr;
  int rc = zkfl_groth16_full_prove_batch(g_ctx, h->c, z, h->r, in, rs, B, p, q, nullptr);
  if (rc) return throw_zkfl(env, rc);
  return make_result(env, proofs, pubs);
}
struct Vk { const uint8_t *alpha1, *beta2, *gamma2, *delta2, *ic; uint32_t n_public; };
bool get_vk(napi_env env, napi_value obj, Vk* vk) {
  size_t l1, l2, l3, l4, l5;
  if (!named_buffer(env, obj, "alpha1", &vk->alpha1, &l1) || !named_buffer(env, obj, "beta2", &vk->beta2, &l2) ||
      !named_buffer(env, obj, "gamma2", &vk->gamma2, &l3) || !named_buffer(env, obj, "delta2", &vk->delta2, &l4) ||
      !named_buffer(env, obj, "ic", &vk->ic, &l5))
    return false;
  napi_value n;
  if (napi_get_named_property(env, obj, "nPublic", &n) != napi_ok || napi_get_value_uint32(env, n, &vk->n_public) != napi_ok) {
    throw_type(env, "verification key: nPublic missing");
    return false;
  }
  if (l1 != 64 || l2 != 128 || l3 != 128 || l4 != 128 || l5 != (size_t)(vk->n_public + 1) * 64) {
    throw_type(env, "verification key: wrong field sizes");
    return false;
  }
  return true;
}
// verify(vk, publics, proof) -> boolean (host pairing check, like one `snarkjs groth16 verify` process)
napi_value Verify(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  Vk vk; const uint8_t *pub, *proof; size_t publen, prooflen;
  if (!get_args(env, info, 3, &argc, argv) || !get_vk(env, argv[0], &vk) || !get_buffer(env, argv[1], &pub, &publen) ||
      !get_buffer(env, argv[2], &proof, &prooflen))
    return nullptr;
  napi_value res;
  if (publen != (size_t)vk.n_public * 32 || prooflen != 256) { napi_get_boolean(env, false, &res); return res; }
  int ok = 0;
  int rc = zkfl_groth16_verify(vk.alpha1, vk.beta2, vk.gamma2, vk.delta2, vk.ic, pub, vk.n_public, proof, &ok);
  if (rc) return throw_zkfl(env, rc);
  napi_get_boolean(env, ok == 1, &res);
  return res;
}
// verifyBatch(vk, publics, proofs, B) -> boolean[]: all proofs of a phase in one GPU pass (Server.verify*Proof, :848-1131)
napi_value VerifyBatch(napi_env env, napi_callback_info info) {
  size_t argc = 4; napi_value argv[4];
  Vk vk; const uint8_t *pub, *proofs; size_t publen, prooflen; int32_t B;
  if (!get_args(env, info, 4, &argc, argv) || !get_vk(env, argv[0], &vk) || !get_buffer(env, argv[1], &pub, &publen) ||
      !get_buffer(env, argv[2], &proofs, &prooflen) || !get_int(env, argv[3], &B) || !ensure_ctx(env))
    return nullptr;
  if (B < 0 || publen != (size_t)B * vk.n_public * 32 || prooflen != (size_t)B * 256) return throw_type(env, "publics / proofs sizes do not match B");
  std::vector<int32_t> ok((size_t)B, 0);
  int rc = zkfl_groth16_verify_batch(g_ctx, vk.alpha1, vk.beta2, vk.gamma2, vk.delta2, vk.ic, vk.n_public, pub, proofs, B, ok.data());
  if (rc) return throw_zkfl(env, rc);
  napi_value arr;
  napi_create_array_with_length(env, (size_t)B, &arr);
  for (int32_t b = 0; b < B; b++) {
    napi_value v;
    napi_get_boolean(env, ok[(size_t)b] == 1, &v);
    napi_set_element(env, arr, (uint32_t)b, v);
  }
  return arr;
}
napi_value ProofToJson(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1];
  const uint8_t* p; size_t len;
  if (!get_args(env, info, 1, &argc, argv) || !get_buffer(env, argv[0], &p, &len)) return nullptr;
  if (len != 256) return throw_type(env, "a proof is 256 bytes");
  std::string buf(2048, '\0');
  int rc = zkfl_proof_to_json(p, &buf[0], buf.size());
  if (rc) return throw_zkfl(env, rc);
  napi_value s;
  napi_create_string_utf8(env, buf.c_str(), NAPI_AUTO_LENGTH, &s);
  return s;
}
napi_value PublicToJson(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2];
  const uint8_t* p; size_t len; int32_t n;
  if (!get_args(env, info, 2, &argc, argv) || !get_buffer(env, argv[0], &p, &len) || !get_int(env, argv[1], &n)) return nullptr;
  if (n < 0 || len != (size_t)n * 32) return throw_type(env, "publics must hold nPublic * 32 bytes");
  std::string buf(84 * (size_t)n + 16, '\0');
  int rc = zkfl_public_to_json(p, (uint32_t)n, &buf[0], buf.size());
  if (rc) return throw_zkfl(env, rc);
  napi_value s;
  napi_create_string_utf8(env, buf.c_str(), NAPI_AUTO_LENGTH, &s);
  return s;
}

napi_value Init(napi_env env, napi_value exports) {
  const napi_property_descriptor props[] = {
      {"loadCircuit", nullptr, LoadCircuit, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"loadZkey", nullptr, LoadZkey, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"circuitInfo", nullptr, CircuitInfo, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"zkeyInfo", nullptr, ZkeyInfo, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"wtnsCalculateBatch", nullptr, WtnsCalculateBatch, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"proveBatch", nullptr, ProveBatch, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"fullProveBatch", nullptr, FullProveBatch, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"verify", nullptr, Verify, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"verifyBatch", nullptr, VerifyBatch, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"proofToJson", nullptr, ProofToJson, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"publicToJson", nullptr, PublicToJson, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
  return exports;
}

}  // namespace

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
