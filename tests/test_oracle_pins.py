"""Pins the oracle: known answers, the reference's only fixture (data/test_input_v5.json, committed as
tests/golden/test_input_v5.json with the script that copied it), and Python <-> C++ agreement."""
import json
import os
import random

import bn254_ref as bn
import groth16_ref as g16
import witness_ref as wr
from parity_cases import tiny_circuit, tiny_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_poseidon_known_answers():
    kats = json.load(open(os.path.join(GOLDEN, "poseidon_kats.json")))
    for k in kats:
        assert bn.poseidon(k["inputs"]) == int(k["hash"])


def test_product_poseidon_parameters_match_oracle():
    import zkfl_b200  # noqa: F401
    from zkfl_b200.circuits import poseidon_params as pp
    for t in (2, 3, 4, 5, 6, 9, 17):
        C, M = bn.poseidon_constants(t)
        C2, M2 = pp.poseidon_params(t)
        assert C == C2 and M == M2
    assert pp.poseidon_hash([1, 2]) == bn.poseidon([1, 2])


def test_fixture_v5_commitments_and_merkle_paths():
    d = json.load(open(os.path.join(GOLDEN, "test_input_v5.json")))
    g = [(int(a) - int(b)) % bn.R for a, b in zip(d["gradPos"], d["gradNeg"])]
    assert bn.gradient_commitment(g, int(d["client_id"]), int(d["round"])) == int(d["root_G"])
    for i in range(8):
        h = bn.vector_hash([int(x) for x in d["features"][i]] + [int(d["labels"][i])])
        for s, p in zip(d["siblings"][i], d["pathIndices"][i]):
            h = bn.poseidon([int(s), h]) if int(p) else bn.poseidon([h, int(s)])
        assert h == int(d["root_D"])


def test_fixture_v5_dataset_regenerates_from_seed():
    import math
    d = json.load(open(os.path.join(GOLDEN, "test_input_v5.json")))
    lcg = bn.JsLcg(42)
    feats, labels = [], []
    for _ in range(128):
        feats.append([math.floor(lcg.random() * 1000) for _ in range(16)])
        labels.append(1 if lcg.random() > 0.5 else 0)
    assert [[str(x) for x in r] for r in feats[:8]] == d["features"]
    assert [str(x) for x in labels[:8]] == d["labels"]
    tree = bn.build_merkle_tree([bn.vector_hash(f + [l]) for f, l in zip(feats, labels)], 7)
    assert tree[-1][0] == int(d["root_D"])
    sib, path = bn.get_merkle_proof(tree, 3, 7)
    assert [str(s) for s in sib] == d["siblings"][3] and [str(p) for p in path] == d["pathIndices"][3]


def test_pairing_bilinear_and_nondegenerate():
    e1 = bn.pairing(bn.G2_GEN, bn.G1_GEN)
    assert not (e1 == bn.Fq12.one())
    assert e1.pow(35) == bn.pairing(bn.ec_mul(bn.G2_GEN, 5), bn.ec_mul(bn.G1_GEN, 7))
    assert bn.ec_mul(bn.G1_GEN, bn.R) is None and bn.ec_mul(bn.G2_GEN, bn.R) is None


def test_cpp_oracle_matches_python_reference(oracle):
    rnd = random.Random(1)
    for _ in range(100):
        a, b = rnd.randrange(bn.R), rnd.randrange(bn.R)
        assert oracle.fr_mul(a, b) == a * b % bn.R
        a, b = rnd.randrange(bn.Q), rnd.randrange(bn.Q)
        assert oracle.fq_mul(a, b) == a * b % bn.Q
    v = [rnd.randrange(bn.R) for _ in range(64)]
    assert oracle.ntt(v) == g16.ntt(v) and oracle.ntt(v, True) == g16.ntt(v, True)
    ks = [rnd.randrange(bn.R) for _ in range(24)]
    bases = oracle.g1_mul_gen(oracle.fes(ks))
    pts = [g16.g1_from_bytes(bases[64 * i:64 * i + 64]) for i in range(24)]
    assert pts[3] == bn.ec_mul(bn.G1_GEN, ks[3])
    sc = [rnd.randrange(bn.R) for _ in range(24)]
    sc[:3] = [0, 1, bn.R - 1]
    ref = bn.ec_msm(pts, sc)
    assert oracle.g1_msm(bases, oracle.fes(sc)) == ref[0].a.to_bytes(32, "little") + ref[1].a.to_bytes(32, "little")
    b2 = oracle.g2_mul_gen(oracle.fes(ks[:8]))
    pts2 = [g16.g2_from_bytes(b2[128 * i:128 * i + 128]) for i in range(8)]
    ref2 = bn.ec_msm(pts2, sc[:8])
    assert oracle.g2_msm(b2, oracle.fes(sc[:8])) == b"".join(
        x.to_bytes(32, "little") for x in (ref2[0].a, ref2[0].b, ref2[1].a, ref2[1].b))


def test_groth16_python_reference_end_to_end_and_cpp_parity(oracle):
    cc = tiny_circuit()
    prog = wr.Program(cc.program_bytes())
    r1 = wr.R1cs(cc.r1cs_bytes())
    inp = tiny_inputs()[0]
    w = wr.calculate_witness(prog, cc.flatten_input(inp))
    assert r1.first_violation(w) is None
    assert oracle.ints(oracle.witness_batch(cc.program_bytes(), oracle.fes(cc.flatten_input(inp)), cc.n_inputs, cc.n_wires)) == w
    assert oracle.poseidon(cc.program_bytes(), [3, 5]) == bn.poseidon([3, 5])
    # python setup is slow (pure-Python scalar muls): use the small golden zkey made by the same function
    zk = open(os.path.join(GOLDEN, "tiny.zkey"), "rb").read()
    proof, pub = g16.prove(zk, w, r=11, s=22)
    vk = g16.Zkey(zk).vkey()
    assert g16.verify(vk, pub, proof)
    assert not g16.verify(vk, [pub[0] + 1, pub[1]], proof)
    pb, pubb = oracle.groth16_prove(zk, oracle.fes(w), 11, 22)
    assert pb == g16.proof_to_bytes(proof) and oracle.ints(pubb) == pub
    assert pb == open(os.path.join(GOLDEN, "tiny.proof"), "rb").read()
    w_bad = list(w)
    w_bad[3] = 4
    proof2, pub2 = g16.prove(zk, w_bad, r=11, s=22)
    assert not g16.verify(vk, pub2, proof2)


def test_final_exponentiation_identities_used_by_the_batch_verifier():
    """csrc/pairing.cuh avoids the Fq12 inversion and the 2816-bit power: f^((p^12-1)/r) == 1 <=> (frob2(conj f) conj f)^h ==
    (frob2(f) f)^h with h = (p^4 - p^2 + 1)/r, conj = odd coefficients negated, frob2 = coefficient i times zeta^i.
    The constants in that header and both maps are checked here against plain powers in the oracle's Fq12."""
    import random
    import re
    Q, R = bn.Q, bn.R
    src = open(os.path.join(os.path.dirname(__file__), "..", "verifiable-federated-training-with-zero-knowledge-proofs-zk-fl-_b200", "csrc", "pairing.cuh")).read()

    def words(name):
        body = re.search(name + r"\[\d+\]\s*=\s*\{([^}]*)\}", src).group(1)
        return sum(int(w.strip().rstrip("u"), 16) << (32 * i) for i, w in enumerate(body.split(",")))
    zeta, h = words("ZETA"), words("H")
    w = bn.Fq12([0, 1] + [0] * 10)
    assert w.pow(Q * Q - 1) == bn.Fq12([zeta] + [0] * 11)            # zeta = xi^((p^2-1)/6) lies in Fq
    assert w.pow(Q ** 6 - 1) == bn.Fq12([Q - 1] + [0] * 11)          # conj: w -> -w
    assert (Q ** 4 - Q * Q + 1) % R == 0 and h == (Q ** 4 - Q * Q + 1) // R
    assert (Q ** 12 - 1) // R == (Q ** 6 - 1) * (Q * Q + 1) * h
    rnd = random.Random(12)
    x = bn.Fq12([rnd.randrange(Q) for _ in range(12)])
    assert bn.Fq12([c * pow(zeta, i, Q) for i, c in enumerate(x.c)]) == x.pow(Q * Q)
    assert bn.Fq12([c if i % 2 == 0 else -c for i, c in enumerate(x.c)]) == x.pow(Q ** 6)


def test_hard_part_chain_of_the_verifier_is_a_multiple_of_h():
    """pairing.cuh's final_exp_is_one raises the easy-part value g to E through conj / squarings / three powers of -x /
    Frobenius^1,2,3; replayed here on exponents modulo the cyclotomic order p^4 - p^2 + 1 = r h: E must be k h with r not
    dividing k, so that g^E == 1 exactly when g^h == 1."""
    p, r = bn.Q, bn.R
    x = 0x44e992b44a6909f1
    assert 36 * x ** 4 + 36 * x ** 3 + 24 * x ** 2 + 6 * x + 1 == p and 36 * x ** 4 + 36 * x ** 3 + 18 * x ** 2 + 6 * x + 1 == r
    phi = p ** 4 - p * p + 1
    h = phi // r

    def negx(e):
        return -e * x % phi
    g = 1
    y0 = negx(g); y1 = 2 * y0; y2 = 2 * y1; y3 = y2 + y1; y4 = negx(y3); y6 = negx(2 * y4)
    y3, y6 = -y3, -y6
    y8 = y6 + y4 + y3
    y9 = y8 + y1
    y11 = y8 + y4 + g
    e = (y9 * p + y11 + y8 * p * p + (y9 - g) * p ** 3) % phi
    assert e % h == 0 and (e // h) % r != 0
