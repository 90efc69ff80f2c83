// zkfl_snarkjs.mjs -- the snarkjs names the reference's tests use, over the zkfl_napi addon (libzkfl.so on a B200).
//
// Drop-in for `import * as snarkjs from "snarkjs"` at the call sites of the reference:
//   groth16.fullProve(input, wasmFile, zkeyFile) -> {proof, publicSignals}    (north star; replaces the witness + prove child
//                                                                             processes of tests/full_system_simulation.mjs:760-775)
//   groth16.prove(zkeyFile, wtnsFile)            -> {proof, publicSignals}    (:773-775)
//   groth16.verify(vkey, publicSignals, proof)   -> boolean                   (:865-868)
//   wtns.calculate(input, wasmFile, wtnsFile)                                 (:760-762, test_secureagg.cjs:108-118)
// plus the batch forms the GPU wants (one call per phase of the round instead of one per client):
//   groth16.fullProveBatch(inputs[], wasmFile, zkeyFile) -> [{proof, publicSignals}, ...]
//   groth16.verifyBatch(vkey, [[publicSignals, proof], ...]) -> boolean[]
// `wasmFile` is the compiled witness program written by `python -m zkfl_b200.cli circom ...` to <name>_js/<name>.wasm (magic
// "zkwp"); its section 8 holds the input map (names, shapes, declaration order) that circom's input.json semantics need.
// Not executed in this repository (no Node.js on the development image); the Python host zkfl_b200/snarkjs.py is the tested twin.
import fs from "node:fs";
import path from "node:path";
import { createRequire } from "node:module";

const require = createRequire(import.meta.url);
const zkfl = require("./build/Release/zkfl_napi.node");

const R = 21888242871839275222246405745257275088548364400416034343698204186575808495617n;
const circuits = new Map();   // file key -> {handle, meta, info}
const zkeys = new Map();

const fileKey = (p) => { const st = fs.statSync(p); return `${path.resolve(p)}:${st.mtimeMs}:${st.size}`; };

function readSections(buf, magic) {       // iden3 binfile container: magic, version u32, nSections u32, (id u32, len u64, bytes)*
  if (buf.toString("latin1", 0, 4) !== magic) throw new Error(`not a ${magic} file`);
  const n = buf.readUInt32LE(8);
  const out = new Map();
  let pos = 12;
  for (let i = 0; i < n; i++) {
    const id = buf.readUInt32LE(pos), len = Number(buf.readBigUInt64LE(pos + 4));
    pos += 12;
    if (!out.has(id)) out.set(id, buf.subarray(pos, pos + len));
    pos += len;
  }
  return out;
}

function loadCircuit(wasmFile) {
  const key = fileKey(wasmFile);
  if (!circuits.has(key)) {
    const prog = fs.readFileSync(wasmFile);
    const base = path.basename(wasmFile).replace(/\.wasm$/, "");
    const r1csFile = path.join(path.dirname(path.dirname(path.resolve(wasmFile))), `${base}.r1cs`);
    const r1cs = fs.existsSync(r1csFile) ? fs.readFileSync(r1csFile) : undefined;   // fullProve refuses to run without it
    const handle = zkfl.loadCircuit(prog, r1cs);
    const meta = JSON.parse(readSections(prog, "zkwp").get(8).toString("utf8"));
    circuits.set(key, { handle, meta, info: zkfl.circuitInfo(handle) });
  }
  return circuits.get(key);
}
function loadZkey(zkeyFile) {
  const key = fileKey(zkeyFile);
  if (!zkeys.has(key)) {
    const handle = zkfl.loadZkey(fs.readFileSync(zkeyFile));
    zkeys.set(key, { handle, info: zkfl.zkeyInfo(handle) });
  }
  return zkeys.get(key);
}

const fe = (v) => {                       // circom input semantics: decimal strings / numbers / bigints, negatives wrap mod r
  let x = BigInt(v) % R;
  if (x < 0n) x += R;
  const b = Buffer.alloc(32);
  for (let i = 0; i < 32; i++) { b[i] = Number(x & 0xffn); x >>= 8n; }
  return b;
};
function flatten(input, meta) {           // input-declaration order of the circuit's main component
  const out = [];
  for (const { name, shape } of meta.inputs) {
    if (!(name in input)) throw new Error(`Signal not found: ${name}`);
    const walk = (v, dims) => {
      if (dims.length === 0) { if (Array.isArray(v)) throw new Error(`Too many values for input signal ${name}`); out.push(fe(v)); return; }
      if (!Array.isArray(v) || v.length !== dims[0]) throw new Error(`Wrong dimensions for input signal ${name}`);
      for (const x of v) walk(x, dims.slice(1));
    };
    walk(input[name], shape);
  }
  return Buffer.concat(out);
}
const dec = (buf, off) => { let x = 0n; for (let i = 31; i >= 0; i--) x = (x << 8n) | BigInt(buf[off + i]); return x.toString(); };
const g1Bytes = (p) => (p[2] === "0" ? Buffer.alloc(64) : Buffer.concat([fe(p[0]), fe(p[1])]));
const g2Bytes = (p) => (p[2][0] === "0" && p[2][1] === "0" ? Buffer.alloc(128) : Buffer.concat([fe(p[0][0]), fe(p[0][1]), fe(p[1][0]), fe(p[1][1])]));
const vkeyBytes = (vk) => ({ alpha1: g1Bytes(vk.vk_alpha_1), beta2: g2Bytes(vk.vk_beta_2), gamma2: g2Bytes(vk.vk_gamma_2),
  delta2: g2Bytes(vk.vk_delta_2), ic: Buffer.concat(vk.IC.map(g1Bytes)), nPublic: Number(vk.nPublic) });
function proofBytes(p) {
  if ((p.protocol ?? "groth16") !== "groth16" || !["bn128", "bn254"].includes(p.curve ?? "bn128")) throw new Error("not a groth16 proof over bn128");
  if (p.pi_a[2] !== "1" || p.pi_c[2] !== "1" || p.pi_b[2][0] !== "1" || p.pi_b[2][1] !== "0") throw new Error("proof points must be affine (z = 1)");
  return Buffer.concat([p.pi_a[0], p.pi_a[1], p.pi_b[0][0], p.pi_b[0][1], p.pi_b[1][0], p.pi_b[1][1], p.pi_c[0], p.pi_c[1]].map(fe));
}
const signalsBytes = (s) => Buffer.concat(s.map(fe));
const unpack = ({ proofs, publics }, B, nPublic) => Array.from({ length: B }, (_, b) => ({
  proof: JSON.parse(zkfl.proofToJson(proofs.subarray(256 * b, 256 * (b + 1)))),
  publicSignals: Array.from({ length: nPublic }, (_, j) => dec(publics, 32 * (nPublic * b + j))),
}));

function wtnsFile(values) {               // .wtns v2: section 1 = n8, prime, nWitness; section 2 = values
  const n = values.length / 32;
  const hdr = Buffer.alloc(40); hdr.writeUInt32LE(32, 0); fe(R - 0n).copy(hdr, 4); fe(0).copy(hdr, 4);
  let r = R; for (let i = 0; i < 32; i++) { hdr[4 + i] = Number(r & 0xffn); r >>= 8n; }
  hdr.writeUInt32LE(n, 36);
  const head = Buffer.alloc(12); head.write("wtns", 0, "latin1"); head.writeUInt32LE(2, 4); head.writeUInt32LE(2, 8);
  const sec = (id, body) => { const h = Buffer.alloc(12); h.writeUInt32LE(id, 0); h.writeBigUInt64LE(BigInt(body.length), 4); return Buffer.concat([h, body]); };
  return Buffer.concat([head, sec(1, hdr), sec(2, values)]);
}

export const wtns = {
  async calculate(input, wasmFile, wtnsOut) {
    const c = loadCircuit(wasmFile);
    const w = zkfl.wtnsCalculateBatch(c.handle, flatten(input, c.meta), 1);   // throws (code ZKFL_ASSERT) on a failed ===
    if (wtnsOut) fs.writeFileSync(wtnsOut, wtnsFile(w));
    return w;
  },
};

export const groth16 = {
  async fullProve(input, wasmFile, zkeyFile) {
    return (await groth16.fullProveBatch([input], wasmFile, zkeyFile))[0];
  },
  async fullProveBatch(inputs, wasmFile, zkeyFile) {
    const c = loadCircuit(wasmFile), z = loadZkey(zkeyFile);
    const packed = Buffer.concat(inputs.map((i) => flatten(i, c.meta)));
    return unpack(zkfl.fullProveBatch(c.handle, z.handle, packed, inputs.length), inputs.length, z.info.nPublic);
  },
  async prove(zkeyFile, wtnsFileOrBuffer) {
    const z = loadZkey(zkeyFile);
    const raw = Buffer.isBuffer(wtnsFileOrBuffer) ? wtnsFileOrBuffer : fs.readFileSync(wtnsFileOrBuffer);
    const values = raw.toString("latin1", 0, 4) === "wtns" ? readSections(raw, "wtns").get(2) : raw;
    return unpack(zkfl.proveBatch(z.handle, values, 1), 1, z.info.nPublic)[0];
  },
  async verify(vkey, publicSignals, proof) {
    const vk = vkeyBytes(vkey);
    if (publicSignals.length !== vk.nPublic) return false;
    let pb; try { pb = proofBytes(proof); } catch { return false; }
    return zkfl.verify(vk, signalsBytes(publicSignals), pb);
  },
  async verifyBatch(vkey, items) {
    const vk = vkeyBytes(vkey);
    const res = new Array(items.length).fill(false), slots = [], pubs = [], proofs = [];
    items.forEach(([s, p], k) => {
      try { if (s.length !== vk.nPublic) throw new Error("count"); pubs.push(signalsBytes(s)); proofs.push(proofBytes(p)); slots.push(k); } catch { /* stays false */ }
    });
    if (slots.length) zkfl.verifyBatch(vk, Buffer.concat(pubs), Buffer.concat(proofs), slots.length).forEach((ok, i) => { res[slots[i]] = ok; });
    return res;
  },
};
