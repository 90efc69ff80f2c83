"""Parity cases shared by the CPU suite (host emulation of the kernels) and the GPU suite (real library).
Every case drives the C ABI through zkfl_b200.api.Prover and compares with the oracle bit for bit."""
from __future__ import annotations

import random

import bn254_ref as bn
import groth16_ref as g16
import oracle_lib as ol
import witness_ref as wr
from zkfl_b200 import inputs as I
from zkfl_b200.circuits import CircuitBuilder, build_circuit
from zkfl_b200.circuits import templates as T
from zkfl_b200.formats import export_verification_key


def tiny_circuit():
    c = CircuitBuilder("tiny")
    out = c.input("out", public=True)
    bound = c.input("bound", public=True)
    x = c.input("x")
    y = c.input("y")
    z = c.mul(x, y)
    c.assert_eq(out, c.mul(z, z, add=x))
    c.assert_eq(T.less_than(c, 8, x, bound), 1)
    c.poseidon([x, y])
    return c.compile()


def tiny_inputs():
    return [{"out": str((3 * 5) ** 2 + 3), "bound": "17", "x": "3", "y": "5"},
            {"out": str((4 * 9) ** 2 + 4), "bound": "200", "x": "4", "y": "9"},
            {"out": "6", "bound": "3", "x": "2", "y": "-1"}]


def case_generator_mul(P, n1=24, n2=8, seed=1):
    rnd = random.Random(seed)
    ks = [rnd.randrange(bn.R) for _ in range(n1)] + [0, 1, bn.R - 1]
    assert P.g1_mul_generator(ks) == ol.g1_mul_gen(ol.fes(ks))
    ks2 = ks[:n2] + [0, 1, bn.R - 1]
    assert P.g2_mul_generator(ks2) == ol.g2_mul_gen(ol.fes(ks2))


def edge_scalars(rnd, n):
    sc = [rnd.randrange(bn.R) for _ in range(n)]
    special = [0, 1, 2, bn.R - 1, bn.R - 2, 1 << 253, (1 << 253) - 1, 0xFFFF, 0x10000, 0x8000, 0x7FFF]
    for i, v in enumerate(special[:n]):
        sc[i] = v
    return sc


def case_g1_msm(P, n, seed=2, bases=None):
    rnd = random.Random(seed)
    if bases is None:
        bases = ol.g1_mul_gen(ol.fes([rnd.randrange(bn.R) for _ in range(n)]))
    sc = ol.fes(edge_scalars(rnd, n))
    assert P.g1_msm(bases, sc) == ol.g1_msm(bases, sc)


def case_g1_msm_degenerate(P):
    """zero scalars, repeated points (doubling inside a bucket), P and -P in one bucket, infinity bases."""
    rnd = random.Random(5)
    one = ol.g1_mul_gen(ol.fes([7]))
    neg = ol.g1_mul_gen(ol.fes([bn.R - 7]))
    bases = one * 6 + neg * 2 + bytes(64) * 2
    for sc in ([0] * 10, [5] * 10, [5, 5, 5, 5, 5, 5, 5, 5, 9, 9], [1, 2, 3, 4, 5, 6, 7, 8, 9, 10],
               [rnd.randrange(bn.R) for _ in range(10)]):
        assert P.g1_msm(bases, ol.fes(sc)) == ol.g1_msm(bases, ol.fes(sc)), sc


def case_g2_msm(P, n, seed=3):
    rnd = random.Random(seed)
    bases = ol.g2_mul_gen(ol.fes([rnd.randrange(bn.R) for _ in range(n)]))
    sc = ol.fes(edge_scalars(rnd, n))
    assert P.g2_msm(bases, sc) == ol.g2_msm(bases, sc)


def case_msm_resident(P, n, group=1, seed=6):
    """resident base set (zkfl_msm_bases_load: window-shifted table, one bucket set) against the oracle and the one-shot call"""
    import ctypes
    rnd = random.Random(seed)
    ks = ol.fes([rnd.randrange(bn.R) for _ in range(n)])
    bases = ol.g1_mul_gen(ks) if group == 1 else ol.g2_mul_gen(ks)
    sc = ol.fes(edge_scalars(rnd, n))
    h = P.msm_load_bases(bases, group)
    out = ctypes.create_string_buffer(64 * group)
    P.msm_run(h, sc, n, out)
    ref = ol.g1_msm(bases, sc) if group == 1 else ol.g2_msm(bases, sc)
    assert out.raw == ref
    P.msm_run(h, sc[:32 * (n - 1)], n - 1, out)          # a prefix of the base set: falls back to the raw points
    assert out.raw == (ol.g1_msm(bases[:64 * (n - 1)], sc[:32 * (n - 1)]) if group == 1 else ol.g2_msm(bases[:128 * (n - 1)], sc[:32 * (n - 1)]))
    P.msm_free_bases(h)


def case_linearity(P, n=64, seed=4):
    """size-independent property: MSM(a) + MSM(b) == MSM(a + b) (checked through the oracle's group law)."""
    rnd = random.Random(seed)
    bases = ol.g1_mul_gen(ol.fes([rnd.randrange(bn.R) for _ in range(n)]))
    a = [rnd.randrange(bn.R) for _ in range(n)]
    b = [rnd.randrange(bn.R) for _ in range(n)]
    pa, pb = P.g1_msm(bases, ol.fes(a)), P.g1_msm(bases, ol.fes(b))
    pab = P.g1_msm(bases, ol.fes([(x + y) % bn.R for x, y in zip(a, b)]))

    def pt(bts):
        return None if bts == bytes(64) else (bn.Fq1(int.from_bytes(bts[:32], "little")), bn.Fq1(int.from_bytes(bts[32:], "little")))
    assert bn.ec_add(pt(pa), pt(pb)) == pt(pab)


def case_witness(P, cc, ins, expect_assert_on=None):
    circ = P.load_circuit(cc)
    ws = P.calculate_witness(circ, ins)
    ref = ol.witness_batch(cc.program_bytes(), circ.pack_inputs(ins), cc.n_inputs, cc.n_wires)
    assert b"".join(ws) == ref
    circ.close()
    return ws


def case_prove(P, cc, ins, rs, seed=b"zkfl-test", python_verify=1):
    """setup on the device, witness + prove through the C ABI, compare with the C++ oracle, verify with the
    oracle's pairing. Returns (zkey bytes, proofs, publics)."""
    circ = P.load_circuit(cc)
    zk = P.new_zkey(cc.r1cs_bytes(), seed)
    Z = P.load_zkey(zk)
    ws = P.calculate_witness(circ, ins)
    proofs, pubs = P.prove(Z, ws, rs)
    for b in range(len(ins)):
        ref_p, ref_pub = ol.groth16_prove(zk, ws[b], *rs[b])
        assert proofs[b] == ref_p, f"proof {b} differs"
        assert pubs[b] == ref_pub
    p2, pub2 = P.full_prove(circ, Z, ins, rs)
    assert p2 == proofs and pub2 == pubs
    vk = g16.vkey_from_json(export_verification_key(zk))
    for b in range(min(python_verify, len(ins))):
        assert g16.verify(vk, ol.ints(pubs[b]), g16.proof_from_bytes(proofs[b]))
    Z.close()
    circ.close()
    return zk, proofs, pubs


def _twist_point_off_subgroup() -> bytes:
    """128 bytes (x.c0, x.c1, y.c0, y.c1 little-endian): a point of the twist y^2 = x^3 + 3/(9+u) that is not in the order-r
    subgroup (the cofactor is ~2^254, so the first point found on the curve is outside it; checked)."""
    q = bn.Q

    def fq2_pow(a, e):
        r = bn.Fq2(1, 0)
        while e:
            if e & 1:
                r = r * a
            a = a * a
            e >>= 1
        return r

    x0 = 1
    while True:
        x = bn.Fq2(x0, 1)
        a = x * x * x + bn.B2
        a1 = fq2_pow(a, (q - 3) // 4)                 # square root in Fq2 for q = 3 mod 4 (Adj, Rodriguez-Henriquez, alg. 9)
        alpha = a1 * a1 * a
        if not (fq2_pow(alpha, q) * alpha == bn.Fq2(q - 1, 0)):
            xr = a1 * a
            y = bn.Fq2(0, 1) * xr if alpha == bn.Fq2(q - 1, 0) else fq2_pow(alpha + bn.Fq2(1, 0), (q - 1) // 2) * xr
            assert y * y == a
            pt, acc, k = (x, y), None, bn.R
            while k:                                 # r * P without reducing the scalar
                if k & 1:
                    acc = bn.ec_add(acc, pt)
                pt = bn.ec_double(pt)
                k >>= 1
            assert acc is not None
            return b"".join(v.to_bytes(32, "little") for v in (x.a, x.b, y.a, y.b))
        x0 += 1


def case_verify_batch(P, zk: bytes, proofs, pubs):
    """GPU (or emulated) batch verifier against the oracle's pairing check and the single-proof host verifier: the valid
    proofs, and a set of tampered / malformed ones."""
    import json
    from zkfl_b200 import formats
    from zkfl_b200 import snarkjs as sj
    vkj = export_verification_key(zk)
    vk = formats.vkey_json_to_bytes(vkj)
    vko = g16.vkey_from_json(vkj)
    n = len(proofs)
    assert P.verify_batch(vk, pubs, proofs) == [True] * n
    psz = len(pubs[0])
    q_le = bn.Q.to_bytes(32, "little")
    bad = []
    bad.append((pubs[0], proofs[1 % n] if n > 1 and proofs[1] != proofs[0] else proofs[0][:64] + proofs[0][64:192] + proofs[0][:64]))  # another proof
    one = (int.from_bytes(pubs[0][:32], "little") + 1) % bn.R
    bad.append((one.to_bytes(32, "little") + pubs[0][32:], proofs[0]))                     # public signal changed
    bad.append((pubs[0], proofs[0][:192] + proofs[0][:64]))                                   # C replaced by A
    bad.append((pubs[0], bytes(64) + proofs[0][64:]))                                         # A = infinity
    bad.append((pubs[0], proofs[0][:32] + (1).to_bytes(32, "little") + proofs[0][64:]))       # A off the curve
    bad.append((pubs[0], q_le + proofs[0][32:]))                                              # coordinate not reduced mod q
    bad.append((bn.R.to_bytes(32, "little") + pubs[0][32:], proofs[0]))                       # public signal not reduced mod r
    bad.append((pubs[0], proofs[0][:64] + bytes(128) + proofs[0][192:]))                      # B = infinity
    assert all(len(b[0]) == psz and len(b[1]) == 256 for b in bad)
    mixed_pubs = [b[0] for b in bad] + list(pubs)
    mixed_proofs = [b[1] for b in bad] + list(proofs)
    got = P.verify_batch(vk, mixed_pubs, mixed_proofs)
    assert got == [False] * len(bad) + [True] * n, got
    # the single-proof host verifier and the oracle agree item by item
    for k in range(len(mixed_proofs)):
        sig, pj = formats.publics_bytes_to_json(mixed_pubs[k]), formats.proof_bytes_to_json(mixed_proofs[k])
        assert sj.groth16.verify(vkj, sig, pj) == got[k], k
    for k in (1, len(bad)):     # pure-Python pairing: slow, two items
        assert g16.verify(vko, ol.ints(mixed_pubs[k]), g16.proof_from_bytes(mixed_proofs[k])) == got[k]
    # ---- the combined (random-linear-combination) check: settles a batch of valid proofs in one final exponentiation ...
    import os
    rlc_on = all(os.environ.get(k, d) == d for k, d in (("ZKFL_VERIFY_FLAT", "0"), ("ZKFL_VERIFY_COOP", "1"), ("ZKFL_VERIFY_RLC", "1")))

    def settled_by_rlc(expect=True):
        import ctypes
        v = ctypes.c_int32(-1)
        P._check(P.lib.zkfl_debug_read(P.ctx, b"v_last_rlc", ctypes.byref(v), 4))
        return (v.value == 1) == expect if rlc_on else True          # the cross-check forms never use the combined check
    assert settled_by_rlc(False)                                           # the mixed batch above fell back to per-proof verdicts
    rep = 2 if n >= 2 else 4
    assert P.verify_batch(vk, list(pubs) * rep, list(proofs) * rep) == [True] * (n * rep)
    assert n * rep < 4 or settled_by_rlc()
    # ... malformed items (A = infinity, coordinate >= q) stay out of the sums and come back False without spoiling the batch
    mal = [bad[3], bad[5]]
    got = P.verify_batch(vk, [b[0] for b in mal] + list(pubs) * rep, [b[1] for b in mal] + list(proofs) * rep)
    assert got == [False, False] + [True] * (n * rep) and settled_by_rlc()
    # ... a B on the twist but OUTSIDE the order-r subgroup makes the weighted sum meaningless: the subgroup test sends the batch
    # to the per-proof form, whose verdict is the single-proof verifier's
    off = _twist_point_off_subgroup()
    odd = proofs[0][:64] + off + proofs[0][192:]
    got = P.verify_batch(vk, [pubs[0]] + list(pubs) * rep, [odd] + list(proofs) * rep)
    assert settled_by_rlc(False)
    assert got == [sj.groth16.verify(vkj, formats.publics_bytes_to_json(pubs[0]), formats.proof_bytes_to_json(odd))] + [True] * (n * rep)
    # ... and the thread-per-proof kernels (ZKFL_VERIFY_COOP=0) give the same verdicts as the lane-cooperative ones
    prev = os.environ.get("ZKFL_VERIFY_COOP")
    os.environ["ZKFL_VERIFY_COOP"] = "0"
    try:
        assert P.verify_batch(vk, mixed_pubs, mixed_proofs) == [False] * len(bad) + [True] * n
    finally:
        if prev is None:
            del os.environ["ZKFL_VERIFY_COOP"]
        else:
            os.environ["ZKFL_VERIFY_COOP"] = prev
    # snarkjs-shaped entry point, including items it cannot even encode
    items = [(formats.publics_bytes_to_json(q), formats.proof_bytes_to_json(p)) for p, q in zip(proofs, pubs)]
    items.append((items[0][0][:-1], items[0][1]))                                             # one signal missing
    assert sj.groth16.verifyBatch(vkj, items, prover=P) == [True] * n + [False]
    assert sj.groth16.verifyBatch(json.loads(json.dumps(vkj)), items[:2], device=False) == [True] * min(2, n)
