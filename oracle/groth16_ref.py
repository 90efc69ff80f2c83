"""ORACLE (test infrastructure). Groth16 over BN254 on Python big ints, small circuits only.

Restates, step for step, what the reference's call sites run inside snarkjs ^0.7.5
(un-vendored dependency, /root/reference/package.json:45):
  * `snarkjs groth16 setup` + `zkey export verificationkey`
    (/root/reference/tests/full_system_simulation.mjs:714-735) -> `setup()`; there is no
    `.ptau` in the tree, so the structured reference string comes from explicit toxic waste.
  * `snarkjs groth16 prove zkey wtns proof.json public.json` (:773-775) -> `prove()`:
    buildABC1, 3x ifft, odd-coset shift by w_{2n}, 3x fft, joinABC, five multiexps,
    blinding with r, s (taken as parameters: snarkjs draws them from a CSPRNG).
  * `snarkjs groth16 verify vkey public proof` (:865-868) -> `verify()`.
`.zkey` layout per SURVEY Appendix A.5.

Parity status: proof bytes are "parity unpinned" against snarkjs itself (no Node.js here,
no zkey/proof fixture in the reference tree).  What is pinned: every proof made here
verifies under the pairing check, and the fast C++ oracle / the CUDA prover must reproduce
these bytes exactly for the same zkey, wtns, r, s.
"""
from __future__ import annotations

import struct

import bn254_ref as bn
from bn254_ref import Fq1, Fq2, G1_GEN, G2_GEN, Q, R, ec_add, ec_mul, ec_neg, fr_root_of_unity
from witness_ref import R1cs, read_sections

RQ = bn.MONT_R_Q
RR = bn.MONT_R_R


# --------------------------------------------------------------------------- byte encodings
def g1_to_bytes(P) -> bytes:
    """affine, little-endian, Montgomery form; infinity = zeros."""
    if P is None:
        return bytes(64)
    return (P[0].a * RQ % Q).to_bytes(32, "little") + (P[1].a * RQ % Q).to_bytes(32, "little")


def g2_to_bytes(P) -> bytes:
    if P is None:
        return bytes(128)
    x, y = P
    return b"".join((v * RQ % Q).to_bytes(32, "little") for v in (x.a, x.b, y.a, y.b))


def g1_from_bytes(b: bytes):
    if b == bytes(64):
        return None
    ri = bn.inv_mod(RQ, Q)
    return (Fq1(int.from_bytes(b[:32], "little") * ri), Fq1(int.from_bytes(b[32:64], "little") * ri))


def g2_from_bytes(b: bytes):
    if b == bytes(128):
        return None
    ri = bn.inv_mod(RQ, Q)
    v = [int.from_bytes(b[32 * i:32 * i + 32], "little") * ri % Q for i in range(4)]
    return (Fq2(v[0], v[1]), Fq2(v[2], v[3]))


# --------------------------------------------------------------------------- NTT
def ntt(vals, inverse=False):
    n = len(vals)
    lg = n.bit_length() - 1
    assert 1 << lg == n
    w = fr_root_of_unity(lg)
    if inverse:
        w = bn.inv_mod(w, R)
    a = list(vals)
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j ^= bit
            bit >>= 1
        j |= bit
        if i < j:
            a[i], a[j] = a[j], a[i]
    length = 2
    while length <= n:
        wl = pow(w, n // length, R)
        for s in range(0, n, length):
            x = 1
            for k in range(length // 2):
                u, v = a[s + k], a[s + k + length // 2] * x % R
                a[s + k] = (u + v) % R
                a[s + k + length // 2] = (u - v) % R
                x = x * wl % R
        length <<= 1
    if inverse:
        ni = bn.inv_mod(n, R)
        a = [x * ni % R for x in a]
    return a


def lagrange_at(tau: int, lg: int):
    """[L_i(tau)] over the size-2^lg domain {w^i}."""
    n = 1 << lg
    w = fr_root_of_unity(lg)
    zt = (pow(tau, n, R) - 1) * bn.inv_mod(n, R) % R
    out, wi = [], 1
    for _ in range(n):
        out.append(zt * wi % R * bn.inv_mod((tau - wi) % R, R) % R)
        wi = wi * w % R
    return out


# --------------------------------------------------------------------------- setup
def setup(r1cs: R1cs, tau: int, alpha: int, beta: int, delta: int):
    """Returns (zkey_bytes, vkey_dict). gamma = 1 as in snarkjs `zkey new`."""
    m, l, nc = r1cs.n_wires, r1cs.n_public, r1cs.n_constraints
    lg = max((nc + l + 1 - 1).bit_length(), 1)
    n = 1 << lg
    L = lagrange_at(tau, lg)
    At, Bt, Ct = [0] * m, [0] * m, [0] * m
    coeffs = []
    for c, (a, b, cc) in enumerate(r1cs.constraints):
        for wire, k in a:
            At[wire] = (At[wire] + k * L[c]) % R
            coeffs.append((0, c, wire, k))
        for wire, k in b:
            Bt[wire] = (Bt[wire] + k * L[c]) % R
            coeffs.append((1, c, wire, k))
        for wire, k in cc:
            Ct[wire] = (Ct[wire] + k * L[c]) % R
    for s in range(l + 1):
        At[s] = (At[s] + L[nc + s]) % R
        coeffs.append((0, nc + s, s, 1))
    dinv = bn.inv_mod(delta, R)
    L2 = lagrange_at(tau, lg + 1)
    g1 = lambda k: ec_mul(G1_GEN, k)  # noqa: E731
    g2 = lambda k: ec_mul(G2_GEN, k)  # noqa: E731
    ic = [g1((beta * At[i] + alpha * Bt[i] + Ct[i]) % R) for i in range(l + 1)]
    pa = [g1(At[i]) for i in range(m)]
    pb1 = [g1(Bt[i]) for i in range(m)]
    pb2 = [g2(Bt[i]) for i in range(m)]
    pc = [g1((beta * At[i] + alpha * Bt[i] + Ct[i]) * dinv % R) for i in range(l + 1, m)]
    ph = [g1(L2[2 * i + 1] * dinv % R) for i in range(n)]
    alpha1, beta1, beta2, gamma2 = g1(alpha), g1(beta), g2(beta), G2_GEN
    delta1, delta2 = g1(delta), g2(delta)

    hdr2 = (struct.pack("<I", 32) + Q.to_bytes(32, "little") + struct.pack("<I", 32)
            + R.to_bytes(32, "little") + struct.pack("<III", m, l, n)
            + g1_to_bytes(alpha1) + g1_to_bytes(beta1) + g2_to_bytes(beta2) + g2_to_bytes(gamma2)
            + g1_to_bytes(delta1) + g2_to_bytes(delta2))
    sec4 = struct.pack("<I", len(coeffs)) + b"".join(
        struct.pack("<III", mtx, c, wire) + (k * RR % R * RR % R).to_bytes(32, "little")
        for mtx, c, wire, k in coeffs)
    sections = [
        (1, struct.pack("<I", 1)), (2, hdr2),
        (3, b"".join(g1_to_bytes(p) for p in ic)), (4, sec4),
        (5, b"".join(g1_to_bytes(p) for p in pa)), (6, b"".join(g1_to_bytes(p) for p in pb1)),
        (7, b"".join(g2_to_bytes(p) for p in pb2)), (8, b"".join(g1_to_bytes(p) for p in pc)),
        (9, b"".join(g1_to_bytes(p) for p in ph)), (10, bytes(64) + struct.pack("<I", 0)),
    ]
    out = [b"zkey", struct.pack("<II", 1, len(sections))]
    for sid, payload in sections:
        out += [struct.pack("<IQ", sid, len(payload)), payload]
    vkey = {"nPublic": l, "alpha1": alpha1, "beta2": beta2, "gamma2": gamma2, "delta2": delta2, "IC": ic}
    return b"".join(out), vkey


# --------------------------------------------------------------------------- zkey reader
class Zkey:
    def __init__(self, data: bytes):
        _, s = read_sections(data, b"zkey")
        assert struct.unpack("<I", s[1])[0] == 1, "not a groth16 zkey"
        h = s[2]
        assert struct.unpack_from("<I", h, 0)[0] == 32 and int.from_bytes(h[4:36], "little") == Q
        assert struct.unpack_from("<I", h, 36)[0] == 32 and int.from_bytes(h[40:72], "little") == R
        self.n_vars, self.n_public, self.domain = struct.unpack_from("<III", h, 72)
        p = 84
        self.alpha1 = g1_from_bytes(h[p:p + 64]); p += 64
        self.beta1 = g1_from_bytes(h[p:p + 64]); p += 64
        self.beta2 = g2_from_bytes(h[p:p + 128]); p += 128
        self.gamma2 = g2_from_bytes(h[p:p + 128]); p += 128
        self.delta1 = g1_from_bytes(h[p:p + 64]); p += 64
        self.delta2 = g2_from_bytes(h[p:p + 128]); p += 128
        self.sections = s

    def g1_section(self, sid):
        b = self.sections[sid]
        return [g1_from_bytes(b[64 * i:64 * i + 64]) for i in range(len(b) // 64)]

    def g2_section(self, sid):
        b = self.sections[sid]
        return [g2_from_bytes(b[128 * i:128 * i + 128]) for i in range(len(b) // 128)]

    def coeffs(self):
        b = self.sections[4]
        n = struct.unpack_from("<I", b, 0)[0]
        r2i = bn.inv_mod(RR * RR % R, R)
        for i in range(n):
            mtx, c, wire = struct.unpack_from("<III", b, 4 + 44 * i)
            yield mtx, c, wire, int.from_bytes(b[16 + 44 * i:48 + 44 * i], "little") * r2i % R

    def vkey(self):
        return {"nPublic": self.n_public, "alpha1": self.alpha1, "beta2": self.beta2,
                "gamma2": self.gamma2, "delta2": self.delta2, "IC": self.g1_section(3)}


# --------------------------------------------------------------------------- prove (snarkjs groth16_prove.js order)
def h_scalars(zk: Zkey, w):
    """Evaluations of A*B - C on the odd coset {w_2n^(2i+1)} (canonical ints)."""
    n = zk.domain
    lg = n.bit_length() - 1
    a_t, b_t = [0] * n, [0] * n
    for mtx, c, wire, k in zk.coeffs():               # buildABC1
        if mtx == 0:
            a_t[c] = (a_t[c] + k * w[wire]) % R
        else:
            b_t[c] = (b_t[c] + k * w[wire]) % R
    c_t = [x * y % R for x, y in zip(a_t, b_t)]
    inc = 25 if lg == 28 else fr_root_of_unity(lg + 1)  # Fr.shift only when power == Fr.s

    def odd(v):
        coef = ntt(v, inverse=True)
        x = 1
        for i in range(n):
            coef[i] = coef[i] * x % R
            x = x * inc % R
        return ntt(coef)

    ao, bo, co = odd(a_t), odd(b_t), odd(c_t)
    return [(x * y - z) % R for x, y, z in zip(ao, bo, co)]   # joinABC


def prove(zkey_bytes: bytes, w, r: int, s: int, msm=bn.ec_msm):
    zk = Zkey(zkey_bytes)
    assert len(w) == zk.n_vars
    l = zk.n_public
    p = h_scalars(zk, w)
    pi_a = msm(zk.g1_section(5), w)
    pi_b1 = msm(zk.g1_section(6), w)
    pi_b = msm(zk.g2_section(7), w)
    pi_c = msm(zk.g1_section(8), w[l + 1:])
    res_h = msm(zk.g1_section(9), p)
    pi_a = ec_add(ec_add(pi_a, zk.alpha1), ec_mul(zk.delta1, r))
    pi_b = ec_add(ec_add(pi_b, zk.beta2), ec_mul(zk.delta2, s))
    pi_b1 = ec_add(ec_add(pi_b1, zk.beta1), ec_mul(zk.delta1, s))
    pi_c = ec_add(pi_c, res_h)
    pi_c = ec_add(pi_c, ec_mul(pi_a, s))
    pi_c = ec_add(pi_c, ec_mul(pi_b1, r))
    pi_c = ec_add(pi_c, ec_mul(zk.delta1, (-(r * s)) % R))
    return {"pi_a": pi_a, "pi_b": pi_b, "pi_c": pi_c}, [w[i] for i in range(1, l + 1)]


def proof_to_bytes(proof) -> bytes:
    """256-byte C-ABI encoding: A (64) | B (128) | C (64), affine canonical little-endian."""
    def g1(P):
        return bytes(64) if P is None else P[0].a.to_bytes(32, "little") + P[1].a.to_bytes(32, "little")

    def g2(P):
        if P is None:
            return bytes(128)
        return b"".join(v.to_bytes(32, "little") for v in (P[0].a, P[0].b, P[1].a, P[1].b))
    return g1(proof["pi_a"]) + g2(proof["pi_b"]) + g1(proof["pi_c"])


def proof_from_bytes(b: bytes):
    def g1(x):
        if x == bytes(64):
            return None
        return (Fq1(int.from_bytes(x[:32], "little")), Fq1(int.from_bytes(x[32:], "little")))

    def g2(x):
        if x == bytes(128):
            return None
        v = [int.from_bytes(x[32 * i:32 * i + 32], "little") for i in range(4)]
        return (Fq2(v[0], v[1]), Fq2(v[2], v[3]))
    return {"pi_a": g1(b[:64]), "pi_b": g2(b[64:192]), "pi_c": g1(b[192:256])}


def proof_to_json(proof) -> dict:
    """snarkjs proof.json shape."""
    a, b, c = proof["pi_a"], proof["pi_b"], proof["pi_c"]
    return {
        "pi_a": [str(a[0].a), str(a[1].a), "1"],
        "pi_b": [[str(b[0].a), str(b[0].b)], [str(b[1].a), str(b[1].b)], ["1", "0"]],
        "pi_c": [str(c[0].a), str(c[1].a), "1"],
        "protocol": "groth16", "curve": "bn128",
    }


def proof_from_json(j: dict):
    a, b, c = j["pi_a"], j["pi_b"], j["pi_c"]
    return {"pi_a": (Fq1(int(a[0])), Fq1(int(a[1]))),
            "pi_b": (Fq2(int(b[0][0]), int(b[0][1])), Fq2(int(b[1][0]), int(b[1][1]))),
            "pi_c": (Fq1(int(c[0])), Fq1(int(c[1])))}


def vkey_from_json(j: dict):
    g1 = lambda p: (Fq1(int(p[0])), Fq1(int(p[1])))  # noqa: E731
    g2 = lambda p: (Fq2(int(p[0][0]), int(p[0][1])), Fq2(int(p[1][0]), int(p[1][1])))  # noqa: E731
    return {"nPublic": j["nPublic"], "alpha1": g1(j["vk_alpha_1"]), "beta2": g2(j["vk_beta_2"]),
            "gamma2": g2(j["vk_gamma_2"]), "delta2": g2(j["vk_delta_2"]), "IC": [g1(p) for p in j["IC"]]}


# --------------------------------------------------------------------------- verify
def verify(vkey, publics, proof) -> bool:
    """e(A,B) == e(alpha,beta) * e(vk_x,gamma) * e(C,delta); public signals must be < r."""
    if len(publics) != vkey["nPublic"] or any(not (0 <= int(x) < R) for x in publics):
        return False
    a, b, c = proof["pi_a"], proof["pi_b"], proof["pi_c"]
    if not (bn.is_on_curve(a, bn.B1) and bn.is_on_curve(c, bn.B1) and bn.is_on_curve(b, bn.B2)):
        return False
    vk_x = vkey["IC"][0]
    for x, ic in zip(publics, vkey["IC"][1:]):
        vk_x = ec_add(vk_x, ec_mul(ic, int(x)))
    return bn.pairing_product_is_one([
        (ec_neg(a), b), (vkey["alpha1"], vkey["beta2"]), (vk_x, vkey["gamma2"]), (c, vkey["delta2"])])


# --------------------------------------------------------------------------- fast setup (C++ oracle scalar muls)
def setup_fast(r1cs: R1cs, tau: int, alpha: int, beta: int, delta: int, nthreads: int = 0):
    """Same key as `setup()` (same formulas, same section layout) with the k*G multiplications done by the C++ oracle
    (oracle_lib.g1_mul_gen / g2_mul_gen) instead of pure Python -- usable at the real circuit sizes. Returns zkey bytes."""
    import oracle_lib as ol
    m, l, nc = r1cs.n_wires, r1cs.n_public, r1cs.n_constraints
    lg = max((nc + l).bit_length(), 1)
    n = 1 << lg

    def lagrange(lg_):
        nn = 1 << lg_
        w = fr_root_of_unity(lg_)
        pw, x = [], 1
        for _ in range(nn):
            pw.append(x)
            x = x * w % R
        zt = (pow(tau, nn, R) - 1) * bn.inv_mod(nn, R) % R
        den = [(tau - p) % R for p in pw]
        pref, acc = [], 1
        for v in den:
            pref.append(acc)
            acc = acc * v % R
        inv = bn.inv_mod(acc, R)
        out = [0] * nn
        for i in range(nn - 1, -1, -1):
            out[i] = zt * pw[i] % R * (inv * pref[i] % R) % R
            inv = inv * den[i] % R
        return out

    L = lagrange(lg)
    At, Bt, Ct = [0] * m, [0] * m, [0] * m
    coeffs = []
    for c, (a, b, cc) in enumerate(r1cs.constraints):
        for wire, k in a:
            At[wire] = (At[wire] + k * L[c]) % R
            coeffs.append((0, c, wire, k))
        for wire, k in b:
            Bt[wire] = (Bt[wire] + k * L[c]) % R
            coeffs.append((1, c, wire, k))
        for wire, k in cc:
            Ct[wire] = (Ct[wire] + k * L[c]) % R
    for s in range(l + 1):
        At[s] = (At[s] + L[nc + s]) % R
        coeffs.append((0, nc + s, s, 1))
    dinv = bn.inv_mod(delta, R)
    comb = [(beta * a + alpha * b + c) % R for a, b, c in zip(At, Bt, Ct)]
    L2 = lagrange(lg + 1)
    g1s = [alpha, beta, delta] + comb[:l + 1] + At + Bt + [c * dinv % R for c in comb[l + 1:]] + [L2[2 * i + 1] * dinv % R for i in range(n)]
    g1 = ol.g1_mul_gen(ol.fes(g1s), nthreads)
    g2 = ol.g2_mul_gen(ol.fes([beta, 1, delta] + Bt), nthreads)
    pos = [0]

    def take(cnt):
        out = g1[64 * pos[0]:64 * (pos[0] + cnt)]
        pos[0] += cnt
        return out

    alpha1, beta1, delta1 = take(1), take(1), take(1)
    ic, pa, pb1, pc, ph = take(l + 1), take(m), take(m), take(m - l - 1), take(n)
    hdr2 = (struct.pack("<I", 32) + Q.to_bytes(32, "little") + struct.pack("<I", 32) + R.to_bytes(32, "little")
            + struct.pack("<III", m, l, n) + alpha1 + beta1 + g2[:128] + g2[128:256] + delta1 + g2[256:384])
    sec4 = struct.pack("<I", len(coeffs)) + b"".join(
        struct.pack("<III", mtx, c, wire) + (k * RR % R * RR % R).to_bytes(32, "little") for mtx, c, wire, k in coeffs)
    sections = [(1, struct.pack("<I", 1)), (2, hdr2), (3, ic), (4, sec4), (5, pa), (6, pb1), (7, g2[384:]), (8, pc), (9, ph),
                (10, bytes(64) + struct.pack("<I", 0))]
    out = [b"zkey", struct.pack("<II", 1, len(sections))]
    for sid, payload in sections:
        out += [struct.pack("<IQ", sid, len(payload)), payload]
    return b"".join(out)
