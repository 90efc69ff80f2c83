"""ORACLE (test infrastructure). Witness calculation + R1CS check on Python big ints.

Restates what `node generate_witness.cjs <wasm> <input.json> <out.wtns>`
(/root/reference/tests/full_system_simulation.mjs:760-762) computes: every signal of the
circuit from its inputs, aborting when a `===` fails.  It interprets the compiled witness
program (`.zkwp`) with an INDEPENDENT Poseidon (oracle/bn254_ref.py) and checks every R1CS
row parsed back from the `.r1cs` bytes, so it shares no arithmetic with the CUDA evaluator.

Parity status: the witness LAYOUT is ours (no .sym/.wtns exists in the reference tree:
"parity unpinned" for internal signal order); values of the public signals and all
hash outputs are pinned by /root/reference/data/test_input_v5.json.
"""
from __future__ import annotations

import json
import struct

from bn254_ref import POSEIDON_RF, POSEIDON_RP, R, poseidon_constants

OP_LIN, OP_MULADD, OP_BITS, OP_POSEIDON = 1, 2, 3, 4
NONE = 0xFFFFFFFF


def read_sections(data: bytes, magic: bytes):
    """iden3 binfile container."""
    assert data[:4] == magic, (data[:4], magic)
    version, n = struct.unpack_from("<II", data, 4)
    pos = 12
    out = {}
    for _ in range(n):
        sid, ln = struct.unpack_from("<IQ", data, pos)
        pos += 12
        out.setdefault(sid, data[pos:pos + ln])
        pos += ln
    return version, out


def _fr(buf, i):
    return int.from_bytes(buf[32 * i:32 * i + 32], "little")


class Program:
    def __init__(self, data: bytes):
        _, s = read_sections(data, b"zkwp")
        (self.n_wires, self.n_public, self.n_inputs, n_ops, n_lcs, n_terms, n_pos,
         n_widths) = struct.unpack("<8I", s[1])
        self.ops = [struct.unpack_from("<5I", s[2], 20 * i) for i in range(n_ops)]
        self.lc_off = struct.unpack(f"<{n_lcs + 1}I", s[3])
        self.lc_wire = struct.unpack(f"<{n_terms}I", s[4])
        self.lc_coef = [_fr(s[5], i) for i in range(n_terms)]
        self.pos_in = struct.unpack(f"<{n_pos}I", s[6])
        self.meta = json.loads(s[8].decode())
        # the embedded Poseidon constants must equal the oracle's own
        p = 0
        for _ in range(n_widths):
            t, rounds, rp, _pad = struct.unpack_from("<4I", s[7], p)
            p += 16
            C, M = poseidon_constants(t)
            assert rounds == POSEIDON_RF + POSEIDON_RP[t - 2] and rp == POSEIDON_RP[t - 2]
            for c in C:
                assert int.from_bytes(s[7][p:p + 32], "little") == c
                p += 32
            for row in M:
                for m in row:
                    assert int.from_bytes(s[7][p:p + 32], "little") == m
                    p += 32

    def lc(self, k, w):
        acc = 0
        for i in range(self.lc_off[k], self.lc_off[k + 1]):
            acc += self.lc_coef[i] * w[self.lc_wire[i]]
        return acc % R


def calculate_witness(prog: Program, inputs: list[int]) -> list[int]:
    """inputs: flat list of the circuit's input signals (already reduced mod r)."""
    assert len(inputs) == prog.n_inputs
    w = [0] * prog.n_wires
    w[0] = 1
    w[1:1 + prog.n_inputs] = [int(x) % R for x in inputs]
    for op, dst, a, b, c in prog.ops:
        if op == OP_LIN:
            w[dst] = prog.lc(a, w)
        elif op == OP_MULADD:
            v = prog.lc(a, w) * prog.lc(b, w)
            if c != NONE:
                v += prog.lc(c, w)
            w[dst] = v % R
        elif op == OP_BITS:
            v = prog.lc(a, w)
            for i in range(b):
                w[dst + i] = (v >> i) & 1
        elif op == OP_POSEIDON:
            t = a
            C, M = poseidon_constants(t)
            rp = POSEIDON_RP[t - 2]
            st = [0] + [w[prog.pos_in[b + i]] for i in range(t - 1)]
            k = dst
            for r in range(POSEIDON_RF + rp):
                st = [(x + C[r * t + i]) % R for i, x in enumerate(st)]
                full = r < POSEIDON_RF // 2 or r >= POSEIDON_RF // 2 + rp
                for i in (range(t) if full else (0,)):
                    x = st[i]
                    x2 = x * x % R
                    x4 = x2 * x2 % R
                    x5 = x4 * x % R
                    w[k], w[k + 1], w[k + 2] = x2, x4, x5
                    k += 3
                    st[i] = x5
                st = [sum(M[i][j] * st[j] for j in range(t)) % R for i in range(t)]
            w[k] = st[0]
        else:
            raise ValueError(f"bad opcode {op}")
    return w


class R1cs:
    def __init__(self, data: bytes):
        _, s = read_sections(data, b"r1cs")
        h = s[1]
        n8 = struct.unpack_from("<I", h, 0)[0]
        assert n8 == 32 and int.from_bytes(h[4:36], "little") == R
        (self.n_wires, self.n_pub_out, self.n_pub_in, self.n_prv_in, self.n_labels,
         self.n_constraints) = struct.unpack_from("<IIIIQI", h, 36)
        self.n_public = self.n_pub_out + self.n_pub_in
        body = s[2]
        p = 0
        self.constraints = []
        for _ in range(self.n_constraints):
            lcs = []
            for _k in range(3):
                nt = struct.unpack_from("<I", body, p)[0]
                p += 4
                terms = []
                for _t in range(nt):
                    wire = struct.unpack_from("<I", body, p)[0]
                    terms.append((wire, int.from_bytes(body[p + 4:p + 36], "little")))
                    p += 36
                lcs.append(terms)
            self.constraints.append(tuple(lcs))
        assert p == len(body)

    def first_violation(self, w):
        for idx, (a, b, c) in enumerate(self.constraints):
            av = sum(k * w[i] for i, k in a) % R
            bv = sum(k * w[i] for i, k in b) % R
            cv = sum(k * w[i] for i, k in c) % R
            if av * bv % R != cv:
                return idx
        return None


def wtns_bytes(w: list[int]) -> bytes:
    """`.wtns` v2."""
    hdr = struct.pack("<I", 32) + R.to_bytes(32, "little") + struct.pack("<I", len(w))
    body = b"".join(int(x).to_bytes(32, "little") for x in w)
    return (b"wtns" + struct.pack("<II", 2, 2) + struct.pack("<IQ", 1, len(hdr)) + hdr
            + struct.pack("<IQ", 2, len(body)) + body)


def read_wtns(data: bytes) -> list[int]:
    _, s = read_sections(data, b"wtns")
    n8 = struct.unpack_from("<I", s[1], 0)[0]
    assert n8 == 32 and int.from_bytes(s[1][4:36], "little") == R
    n = struct.unpack_from("<I", s[1], 36)[0]
    return [_fr(s[2], i) for i in range(n)]
