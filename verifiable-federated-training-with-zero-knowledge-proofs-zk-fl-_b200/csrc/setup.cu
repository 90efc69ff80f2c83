// libzkfl.so: Groth16 key generation on the GPU -- the `snarkjs groth16 setup <r1cs> <ptau> <zkey>` replacement
// (tests/full_system_simulation.mjs:714-717).  No `.ptau` exists in the reference tree, so the structured
// reference string comes from explicit toxic waste (tau, alpha, beta, delta; gamma = 1 as in snarkjs).  Everything that scales with
// the circuit runs on the device: the Lagrange basis at tau (domain n and the odd half of domain 2n), the per-wire column sums of
// A, B, C against it, the key scalars, and the scalar multiplications of the generators (byte-window tables, 32 mixed additions
// per point, one shared inversion per four points).  The host orders the matrix entries by wire (counting sort) and writes the
// `.zkey` container (snarkjs section layout).
#include "host.h"
#include "k_setup.cuh"

namespace {

struct ByWire {              // one matrix, entries ordered by wire and cut into pieces of at most PIECE entries of one wire
  std::vector<uint32_t> rows, cidx, piece_off, wire_piece;
};
const uint32_t PIECE = 1024;

int order_by_wire(uint32_t m, uint32_t n_rows, uint32_t n_coef, const uint32_t* rows, const uint32_t* wires, const uint32_t* cidx, size_t nnz,
                  ByWire& out) {
  std::vector<uint32_t> off(m + 1, 0);
  for (size_t i = 0; i < nnz; i++) {
    if (rows[i] >= n_rows || wires[i] >= m || cidx[i] >= n_coef) return fail(ZKFL_ERR_ARG, "setup: matrix entry out of range");
    off[wires[i] + 1]++;
  }
  for (uint32_t w = 0; w < m; w++) off[w + 1] += off[w];
  out.rows.resize(nnz); out.cidx.resize(nnz);
  std::vector<uint32_t> cur(off.begin(), off.end() - 1);
  for (size_t i = 0; i < nnz; i++) { const uint32_t p = cur[wires[i]]++; out.rows[p] = rows[i]; out.cidx[p] = cidx[i]; }
  out.wire_piece.assign(m + 1, 0);
  out.piece_off.clear();
  for (uint32_t w = 0; w < m; w++) {
    out.wire_piece[w] = (uint32_t)out.piece_off.size();
    for (uint32_t s = off[w]; s < off[w + 1]; s += PIECE) out.piece_off.push_back(s);
  }
  out.wire_piece[m] = (uint32_t)out.piece_off.size();
  out.piece_off.push_back((uint32_t)nnz);
  // a piece ends where the next piece of the same wire starts, or at the end of the wire: make every piece end explicit
  // (piece p covers [piece_off[p], min(piece_off[p] + PIECE, end of its wire))) by inserting nothing: consecutive starts of one
  // wire are PIECE apart and the first start of the next wire is the end of this one, so piece_off[p + 1] is always the end.
  return 0;
}

Fr fr_from_canonical(const uint8_t* p) { Fr r; memcpy(r.v, p, 32); return r; }

struct Out {                 // bump writer into the caller's buffer
  uint8_t* p; size_t cap, len;
  void put(const void* src, size_t n) { if (p && len + n <= cap) memcpy(p + len, src, n); len += n; }
  void u32(uint32_t v) { put(&v, 4); }
  void u64(uint64_t v) { put(&v, 8); }
  uint8_t* reserve(size_t n) { uint8_t* at = (p && len + n <= cap) ? p + len : nullptr; len += n; return at; }
};

}  // namespace

extern "C" int zkfl_groth16_setup(zkfl_ctx* c, uint32_t n_wires, uint32_t n_public, uint32_t n_constraints, const uint32_t* const rows[3],
                                  const uint32_t* const wires[3], const uint32_t* const cidx[3], const size_t nnz[3],
                                  const uint8_t* coef_table, uint32_t n_coef, const uint8_t toxic[128], uint8_t* zkey_out, size_t cap,
                                  size_t* zkey_len) {
  if (!c || !rows || !wires || !cidx || !nnz || !coef_table || !toxic || !zkey_len || n_wires <= n_public || n_coef == 0)
    return fail(ZKFL_ERR_ARG, "bad argument");
  const uint32_t m = n_wires, l = n_public, nc = n_constraints;
  uint32_t lg = 1;
  while ((1ull << lg) < (uint64_t)nc + l + 1) lg++;      // smallest 2^lg >= nc + l + 1 (the extra rows bind the public inputs)
  if (lg > 27) return fail(ZKFL_ERR_ARG, "setup: domain too large");
  const uint32_t n = 1u << lg;
  for (int k = 0; k < 3; k++) if (nnz[k] > 0xFFFFFFF0ull || (nnz[k] && (!rows[k] || !wires[k] || !cidx[k]))) return fail(ZKFL_ERR_ARG, "bad argument");
  // ---- sizes of the sections; with zkey_out == NULL only the total length is reported
  const size_t n_coef_entries = nnz[0] + nnz[1] + (size_t)l + 1;
  const size_t sec_len[11] = {0, 4, 84 + 64 * 3 + 128 * 3, 64 * ((size_t)l + 1), 4 + 44 * n_coef_entries, 64 * (size_t)m, 64 * (size_t)m,
                              128 * (size_t)m, 64 * (size_t)(m - l - 1), 64 * (size_t)n, 68};
  size_t total = 12;
  for (int s = 1; s <= 10; s++) total += 12 + sec_len[s];
  *zkey_len = total;
  if (!zkey_out) return 0;
  if (cap < total) return fail(ZKFL_ERR_ARG, "setup: output buffer too small");
  for (int k = 0; k < 4; k++) {
    if (!fr_bytes_lt_mod(toxic + 32 * k)) return fail(ZKFL_ERR_ARG, "setup: toxic value not reduced mod r");
    bool zero = true;
    for (int i = 0; i < 32; i++) zero = zero && toxic[32 * k + i] == 0;
    if (zero) return fail(ZKFL_ERR_ARG, "setup: toxic value is zero");
  }
  for (uint32_t i = 0; i < n_coef; i++) if (!fr_bytes_lt_mod(coef_table + 32 * (size_t)i)) return fail(ZKFL_ERR_ARG, "setup: coefficient not reduced mod r");
  CU(cudaSetDevice(c->device));
  const Fr tau = fr_from_canonical(toxic).to_mont(), alpha = fr_from_canonical(toxic + 32).to_mont(),
           beta = fr_from_canonical(toxic + 64).to_mont(), delta = fr_from_canonical(toxic + 96).to_mont();
  // ---- host scalars: tau^n, the two vanishing-polynomial factors, roots
  Fr tn = tau;
  for (uint32_t i = 0; i < lg; i++) tn = tn.sqr();                   // tau^n
  const Fr t2n = tn.sqr();
  if (tn == Fr::one() || t2n == Fr::one()) return fail(ZKFL_ERR_ARG, "setup: tau is a root of unity of the domain");
  Fr nn = Fr::zero(); nn.v[0] = n; nn = nn.to_mont();
  const Fr n_inv = nn.inv(), zt = (tn - Fr::one()) * n_inv, zt2 = (t2n - Fr::one()) * (n_inv * (Fr::one() + Fr::one()).inv());
  const Fr wn = fr_root_of_unity((int)lg), w2n = fr_root_of_unity((int)lg + 1);
  const Fr delta_inv = delta.inv();
  // ---- device: Lagrange bases
  DevBuf dL, dLodd, dCoef, dAt[3], dSc, dSc2, dTab1, dTab2, dPts1, dPts2;
  TRY(dL.reserve((size_t)n * sizeof(Fr))); TRY(dLodd.reserve((size_t)n * sizeof(Fr)));
  {
    Stage st(c, "setup_lagrange");
    ZK_LAUNCH(k_lagrange, ((size_t)n + ZK_LAG_CH - 1) / ZK_LAG_CH, 64, c->stream, tau, Fr::one(), wn, zt, n, dL.as<Fr>());
    ZK_LAUNCH(k_lagrange, ((size_t)n + ZK_LAG_CH - 1) / ZK_LAG_CH, 64, c->stream, tau, w2n, wn, zt2, n, dLodd.as<Fr>());
    CU(cudaGetLastError());
  }
  // ---- column sums At, Bt, Ct
  std::vector<Fr> coef_m(n_coef);
  for (uint32_t i = 0; i < n_coef; i++) coef_m[i] = fr_from_canonical(coef_table + 32 * (size_t)i).to_mont();
  TRY(upload(c, dCoef, coef_m.data(), coef_m.size()));
  for (int k = 0; k < 3; k++) {
    ByWire bw;
    TRY(order_by_wire(m, nc, n_coef, rows[k], wires[k], cidx[k], nnz[k], bw));
    DevBuf dRows, dCidx, dPiece, dWirePiece, dPartial;
    TRY(upload(c, dRows, bw.rows.data(), bw.rows.size())); TRY(upload(c, dCidx, bw.cidx.data(), bw.cidx.size()));
    TRY(upload(c, dPiece, bw.piece_off.data(), bw.piece_off.size())); TRY(upload(c, dWirePiece, bw.wire_piece.data(), bw.wire_piece.size()));
    const size_t n_pieces = bw.piece_off.size() - 1;
    TRY(dPartial.reserve((n_pieces ? n_pieces : 1) * sizeof(Fr)));
    TRY(dAt[k].reserve((size_t)m * sizeof(Fr)));
    Stage st(c, "setup_colsum");
    ZK_LAUNCH(k_colsum_pieces, n_pieces, 128, c->stream, dPiece.as<uint32_t>(), dRows.as<uint32_t>(), dCidx.as<uint32_t>(), dCoef.as<Fr>(),
              dL.as<Fr>(), n_pieces, dPartial.as<Fr>());
    ZK_LAUNCH(k_colsum_wires, m, 128, c->stream, dWirePiece.as<uint32_t>(), dPartial.as<Fr>(), m, dAt[k].as<Fr>());
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));       // the staging buffers of this matrix go out of scope
  }
  // ---- key scalars and the points
  const size_t n_g1 = 3 + ((size_t)l + 1) + 2 * (size_t)m + (m - l - 1) + n, n_g2 = 3 + (size_t)m;
  TRY(dSc.reserve(n_g1 * sizeof(Fr))); TRY(dSc2.reserve(n_g2 * sizeof(Fr)));
  TRY(dTab1.reserve(32 * 256 * sizeof(G1Affine))); TRY(dTab2.reserve(32 * 256 * sizeof(G2Affine)));
  TRY(dPts1.reserve(n_g1 * sizeof(G1Affine))); TRY(dPts2.reserve(n_g2 * sizeof(G2Affine)));
  {
    Stage st(c, "setup_scalars");
    ZK_LAUNCH(k_setup_scalars, (size_t)(m > n ? m : n), 128, c->stream, dAt[0].as<Fr>(), dAt[1].as<Fr>(), dAt[2].as<Fr>(), dL.as<Fr>(),
              dLodd.as<Fr>(), alpha, beta, delta, delta_inv, m, l, nc, n, dSc.as<Fr>(), dSc2.as<Fr>());
    CU(cudaGetLastError());
  }
  TRY(msm_fixed_base_table<Fq>(c, g1_generator(), dTab1.as<G1Affine>()));
  TRY(msm_fixed_base_table<Fq2>(c, g2_generator(), dTab2.as<G2Affine>()));
  {
    Stage st(c, "setup_points");
    ZK_LAUNCH(k_fixed_base_batch<Fq>, (n_g1 + ZK_FB_CH - 1) / ZK_FB_CH, 64, c->stream, (const G1Affine*)dTab1.as<G1Affine>(), (const Fr*)dSc.as<Fr>(),
              n_g1, dPts1.as<G1Affine>());
    ZK_LAUNCH(k_fixed_base_batch<Fq2>, (n_g2 + ZK_FB_CH - 1) / ZK_FB_CH, 64, c->stream, (const G2Affine*)dTab2.as<G2Affine>(), (const Fr*)dSc2.as<Fr>(),
              n_g2, dPts2.as<G2Affine>());
    CU(cudaGetLastError());
  }
  // ---- the .zkey container, written straight into the caller's buffer (point sections are device-to-host copies)
  Out o{zkey_out, cap, 0};
  o.put("zkey", 4); o.u32(1); o.u32(10);
  auto section = [&](uint32_t id) { o.u32(id); o.u64(sec_len[id]); };
  section(1); o.u32(1);                                                       // protocol: groth16
  section(2);
  uint8_t* hdr = o.reserve(sec_len[2]);
  {
    uint32_t v = 32; memcpy(hdr, &v, 4);
    for (int i = 0; i < 8; i++) { uint32_t q = FqP::mod(i), r = FrP::mod(i); memcpy(hdr + 4 + 4 * i, &q, 4); memcpy(hdr + 40 + 4 * i, &r, 4); }
    memcpy(hdr + 36, &v, 4);
    memcpy(hdr + 72, &m, 4); memcpy(hdr + 76, &l, 4); memcpy(hdr + 80, &n, 4);
  }
  const G1Affine* p1 = dPts1.as<G1Affine>();
  const G2Affine* p2 = dPts2.as<G2Affine>();
  auto d2h = [&](uint8_t* dst, const void* src, size_t bytes) -> int {
    if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
  };
  // header points: alpha1, beta1, beta2, gamma2, delta1, delta2
  TRY(d2h(hdr + 84, p1 + 0, 64)); TRY(d2h(hdr + 148, p1 + 1, 64)); TRY(d2h(hdr + 212, p2 + 0, 128)); TRY(d2h(hdr + 340, p2 + 1, 128));
  TRY(d2h(hdr + 468, p1 + 2, 64)); TRY(d2h(hdr + 532, p2 + 2, 128));
  const size_t o_ic = 3, o_a = o_ic + l + 1, o_b = o_a + m, o_c = o_b + m, o_h = o_c + (m - l - 1);
  section(3); TRY(d2h(o.reserve(sec_len[3]), p1 + o_ic, sec_len[3]));
  section(4);
  {
    uint8_t* s4 = o.reserve(sec_len[4]);
    const uint32_t cnt = (uint32_t)n_coef_entries;
    memcpy(s4, &cnt, 4);
    std::vector<Fr> coef_r2(n_coef);                                           // value * R^2 mod r: what snarkjs stores
    for (uint32_t i = 0; i < n_coef; i++) coef_r2[i] = coef_m[i].to_mont();
    uint8_t* e = s4 + 4;
    for (uint32_t k = 0; k < 2; k++)
      for (size_t i = 0; i < nnz[k]; i++, e += 44) {
        memcpy(e, &k, 4); memcpy(e + 4, rows[k] + i, 4); memcpy(e + 8, wires[k] + i, 4); memcpy(e + 12, coef_r2[cidx[k][i]].v, 32);
      }
    const Fr one_r2 = Fr::one().to_mont();
    for (uint32_t s = 0; s <= l; s++, e += 44) {
      const uint32_t zero = 0, row = nc + s;
      memcpy(e, &zero, 4); memcpy(e + 4, &row, 4); memcpy(e + 8, &s, 4); memcpy(e + 12, one_r2.v, 32);
    }
  }
  section(5); TRY(d2h(o.reserve(sec_len[5]), p1 + o_a, sec_len[5]));
  section(6); TRY(d2h(o.reserve(sec_len[6]), p1 + o_b, sec_len[6]));
  section(7); TRY(d2h(o.reserve(sec_len[7]), p2 + 3, sec_len[7]));
  section(8); TRY(d2h(o.reserve(sec_len[8]), p1 + o_c, sec_len[8]));
  section(9); TRY(d2h(o.reserve(sec_len[9]), p1 + o_h, sec_len[9]));
  section(10);
  { uint8_t* s10 = o.reserve(68); memset(s10, 0, 68); }                        // csHash = 0, no contributions yet (zkey_setup.contribute appends)
  CU(cudaStreamSynchronize(c->stream));
  if (o.len != total) return fail(ZKFL_ERR_ARG, "setup: internal size mismatch");
  return 0;
}
