"""`snarkjs groth16 setup <r1cs> <ptau> <zkey>` + `zkey contribute` replacement
(tests/full_system_simulation.mjs:714-731).

No `.ptau` exists in the reference tree and none can be fetched, so the structured reference string
is derived from explicit toxic waste (tau, alpha, beta, delta; gamma = 1 as in snarkjs).  SOUNDNESS: whoever knows
the toxic waste can forge proofs, so outside tests the seed comes from the OS CSPRNG (`Prover.new_zkey(seed=None)`,
`snarkjs.zKey.newZKey`) and is never stored; `contribute` re-randomises delta with fresh secret entropy the way
`snarkjs zkey contribute` does, so the key stays sound as long as ONE contributor (or the initial setup) was honest.  The field
arithmetic that builds the key scalars runs on the host (Python ints, one-off per circuit); every
scalar multiplication k*G1 / k*G2 runs on the GPU (`zkfl_g1_mul_generator`).  Output: a `.zkey` with
the exact snarkjs section layout, including the nPublic+1 extra A rows and the
odd-coset Lagrange H points the snarkjs prover expects.
"""
from __future__ import annotations

import hashlib
import os
import struct

from .formats import FQ, FR, read_container, write_container

_R2 = pow(1 << 256, 2, FR)


def toxic_from_seed(seed: bytes):
    out = []
    ctr = 0
    while len(out) < 4:
        v = int.from_bytes(hashlib.sha512(seed + bytes([ctr])).digest(), "little") % FR
        ctr += 1
        if v > 1:
            out.append(v)
    return tuple(out)  # tau, alpha, beta, delta


def root_of_unity(power: int) -> int:
    w = pow(5, (FR - 1) >> 28, FR)
    for _ in range(28 - power):
        w = w * w % FR
    return w


def _batch_inverse(vals):
    pref, acc = [], 1
    for v in vals:
        pref.append(acc)
        acc = acc * v % FR
    inv = pow(acc, -1, FR)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        out[i] = inv * pref[i] % FR
        inv = inv * vals[i] % FR
    return out


def lagrange_at(tau: int, lg: int):
    n = 1 << lg
    w = root_of_unity(lg)
    pw, x = [], 1
    for _ in range(n):
        pw.append(x)
        x = x * w % FR
    zt = (pow(tau, n, FR) - 1) * pow(n, -1, FR) % FR
    inv = _batch_inverse([(tau - p) % FR for p in pw])
    return [zt * p % FR * i % FR for p, i in zip(pw, inv)]


def parse_r1cs(data: bytes):
    s = read_container(data, b"r1cs")
    h = s[1]
    if struct.unpack_from("<I", h, 0)[0] != 32 or int.from_bytes(h[4:36], "little") != FR:
        raise ValueError("r1cs: not over the BN254 scalar field")
    n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints = struct.unpack_from("<IIIIQI", h, 36)
    body = s[2]
    mats = ([], [], [])  # per matrix: list of (row, wire, coef)
    p = 0
    for row in range(n_constraints):
        for k in range(3):
            nt = struct.unpack_from("<I", body, p)[0]
            p += 4
            for _ in range(nt):
                wire = struct.unpack_from("<I", body, p)[0]
                mats[k].append((row, wire, int.from_bytes(body[p + 4:p + 36], "little")))
                p += 36
    return {"n_wires": n_wires, "n_public": n_pub_out + n_pub_in, "n_constraints": n_constraints,
            "A": mats[0], "B": mats[1], "C": mats[2]}


def r1cs_info(data: bytes) -> dict:
    s = read_container(data, b"r1cs")
    n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints = struct.unpack_from("<IIIIQI", s[1], 36)
    return {"nWires": n_wires, "nConstraints": n_constraints, "nPrvInputs": n_prv_in,
            "nPubInputs": n_pub_in, "nOutputs": n_pub_out, "nLabels": n_labels}


class _Terms:
    """(row, wire, coef) triples of the three matrices, re-iterable without materialising tuples (the scaled circuits
    have tens of millions of non-zeros). Built from `.r1cs` bytes or straight from a CompiledCircuit."""

    def __init__(self, r1cs):
        if isinstance(r1cs, (bytes, bytearray)):
            d = parse_r1cs(r1cs)
            self.n_wires, self.n_public, self.n_constraints = d["n_wires"], d["n_public"], d["n_constraints"]
            self._m = {k: tuple(zip(*d[k])) if d[k] else ((), (), ()) for k in "ABC"}
        else:
            self.n_wires, self.n_public, self.n_constraints = r1cs.n_wires, r1cs.n_public, r1cs.n_constraints
            self._m = {"A": (r1cs.A.rows, r1cs.A.wires, r1cs.A.coefs), "B": (r1cs.B.rows, r1cs.B.wires, r1cs.B.coefs),
                       "C": (r1cs.C.rows, r1cs.C.wires, r1cs.C.coefs)}

    def __getitem__(self, key):
        return zip(*self._m[key])

    def count(self, key):
        return len(self._m[key][0])


def new_zkey(prover, r1cs, seed: bytes) -> bytes:
    """`groth16 setup` on the GPU (zkfl_groth16_setup, csrc/setup.cu): Lagrange basis, column sums, key scalars and every scalar
    multiplication run on the device and the library writes the `.zkey`.  The host only flattens the three matrices into
    coordinate arrays with a table of the distinct coefficients.  r1cs: `.r1cs` bytes or a CompiledCircuit.
    `seed` determines the toxic waste: pass OS randomness (Prover.new_zkey(seed=None) does) unless a test needs a fixed key."""
    import ctypes
    from array import array
    from . import _lib
    r = _Terms(r1cs)
    table, index = [], {}
    arrs = []
    for key in "ABC":
        rows, wires, coefs = r._m[key]
        idx = []
        for cval in coefs:
            k = index.get(cval)
            if k is None:
                k = index[cval] = len(table)
                table.append(cval)
            idx.append(k)
        arrs.append((array("I", rows), array("I", wires), array("I", idx)))
    coef_bytes = b"".join((int(v) % FR).to_bytes(32, "little") for v in table) or bytes(32)
    toxic = b"".join(int(v).to_bytes(32, "little") for v in toxic_from_seed(seed))
    vp = ctypes.c_void_p

    def ptrs(k):
        return (vp * 3)(*[a[k].buffer_info()[0] if len(a[k]) else None for a in arrs])
    rows_p, wires_p, cidx_p = ptrs(0), ptrs(1), ptrs(2)
    nnz = (ctypes.c_size_t * 3)(*[len(a[0]) for a in arrs])
    need = ctypes.c_size_t(0)
    lib = prover.lib
    args = (prover.ctx, r.n_wires, r.n_public, r.n_constraints, rows_p, wires_p, cidx_p, nnz, _lib.as_ptr(coef_bytes), max(len(table), 1),
            _lib.as_ptr(toxic))
    prover._check(lib.zkfl_groth16_setup(*args, None, 0, ctypes.byref(need)))
    out = bytearray(need.value)
    prover._check(lib.zkfl_groth16_setup(*args, _lib.as_ptr(out), len(out), ctypes.byref(need)))
    return bytes(out)


def contribute(prover, zkey: bytes, name: str = "", entropy: bytes = b"") -> bytes:
    """`snarkjs zkey contribute <in> <out> --name= -e=` (tests/full_system_simulation.mjs:726-731): a fresh secret d
    (OS CSPRNG mixed with the caller's entropy) rescales delta1, delta2 by d and the C and H sections by 1/d, so that nobody
    who does not know d knows the new delta; a contribution record is appended to section 10 (deltaAfter, d*G1 data, a hash
    chaining the previous transcript hash, the name).  The record has snarkjs's field layout but is NOT a verifiable ceremony
    transcript (`zkey verify` needs the ptau the reference tree does not have); the scalar multiplications run on the GPU."""
    s = read_container(zkey, b"zkey")
    h = bytearray(s[2])
    d = 0
    while d < 2:
        d = int.from_bytes(hashlib.sha512(os.urandom(64) + entropy + name.encode()).digest(), "little") % FR
    dinv = pow(d, -1, FR)
    o_d1, o_d2 = 84 + 64 + 64 + 128 + 128, 84 + 64 + 64 + 128 + 128 + 64
    h[o_d1:o_d1 + 64] = prover.scale_points(bytes(h[o_d1:o_d1 + 64]), d, 1)
    h[o_d2:o_d2 + 128] = prover.scale_points(bytes(h[o_d2:o_d2 + 128]), d, 2)
    pc = prover.scale_points(s[8], dinv, 1) if len(s[8]) else b""
    ph = prover.scale_points(s[9], dinv, 1)
    prev = s.get(10, bytes(64) + struct.pack("<I", 0))
    cs_hash, n_contrib = prev[:64], struct.unpack_from("<I", prev, 64)[0]
    g1 = prover.g1_mul_generator([d])
    transcript = hashlib.sha512(cs_hash + bytes(h[o_d1:o_d1 + 64]) + name.encode()).digest()
    nm = name.encode()
    params = bytes([1, len(nm)]) + nm if nm else b""        # snarkjs: type 1 = name
    record = bytes(h[o_d1:o_d1 + 64]) + g1 + bytes(64) + bytes(128) + transcript + struct.pack("<I", len(params)) + params
    sec10 = cs_hash + struct.pack("<I", n_contrib + 1) + prev[68:] + record
    sections = [(sid, {2: bytes(h), 8: pc, 9: ph, 10: sec10}.get(sid, s[sid])) for sid in sorted(set(s) | {10})]
    return write_container(b"zkey", 1, sections)
