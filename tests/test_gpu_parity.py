"""GPU parity tests: the real sm_100a library through the C ABI versus the oracle, bit for bit."""
import random

import pytest

import bn254_ref as bn
import oracle_lib as ol
import parity_cases as pc
from zkfl_b200 import _lib
from zkfl_b200 import inputs as I
from zkfl_b200.circuits import build_circuit

pytestmark = pytest.mark.gpu


def test_library_is_the_cuda_build(gpu_prover):
    assert b"sm_100a" in gpu_prover.lib.zkfl_version()
    before = gpu_prover.launch_count()
    gpu_prover.g1_mul_generator([5])
    assert gpu_prover.launch_count() > before


def test_generator_mul(gpu_prover):
    pc.case_generator_mul(gpu_prover, n1=2000, n2=300)


def test_g1_msm_sizes_and_edges(gpu_prover):
    for n in (1, 2, 33, 1000, 1 << 14):
        pc.case_g1_msm(gpu_prover, n)
    pc.case_g1_msm_degenerate(gpu_prover)
    pc.case_linearity(gpu_prover, n=4096)


def test_g1_msm_2pow20_against_oracle(gpu_prover):
    """BASELINE.json's MSM size. Bases k_i*G come from the (separately checked) device generator kernel."""
    n = 1 << 20
    rnd = random.Random(20)
    bases = gpu_prover.g1_mul_generator(b"".join(rnd.randrange(bn.R).to_bytes(32, "little") for _ in range(n)))
    spot = [0, 1, n // 2, n - 1]
    ks = random.Random(20)
    allk = [ks.randrange(bn.R) for _ in range(n)]
    for i in spot:
        assert bases[64 * i:64 * i + 64] == ol.g1_mul_gen(ol.fe(allk[i]))
    sc = b"".join(rnd.randrange(bn.R).to_bytes(32, "little") for _ in range(n))
    assert gpu_prover.g1_msm(bases, sc) == ol.g1_msm(bases, sc)


def test_resident_msm_with_window_table(gpu_prover):
    """zkfl_msm_bases_load builds the window-shifted table (one bucket set per MSM): same result as the oracle, G1 and G2"""
    for n in (1024, 5000, 1 << 16):
        pc.case_msm_resident(gpu_prover, n, 1)
    pc.case_msm_resident(gpu_prover, 3000, 2)
    pc.case_msm_resident(gpu_prover, 100, 1)


def test_g1_msm_2pow20_resident_table_against_oracle(gpu_prover):
    """BASELINE.json's MSM size through the resident path the bench measures (table, c = 17, one set of 2^16 buckets)"""
    import ctypes
    n = 1 << 20
    rnd = random.Random(21)
    bases = gpu_prover.g1_mul_generator(b"".join(rnd.randrange(bn.R).to_bytes(32, "little") for _ in range(n)))
    sc = b"".join(rnd.randrange(bn.R).to_bytes(32, "little") for _ in range(n))
    h = gpu_prover.msm_load_bases(bases, 1)
    out = ctypes.create_string_buffer(64)
    gpu_prover.msm_run(h, sc, n, out)
    gpu_prover.msm_free_bases(h)
    assert out.raw == ol.g1_msm(bases, sc)


def test_g2_msm(gpu_prover):
    for n in (1, 40, 4096):
        pc.case_g2_msm(gpu_prover, n)


def test_tiny_circuit(gpu_prover):
    cc = pc.tiny_circuit()
    pc.case_witness(gpu_prover, cc, pc.tiny_inputs())
    pc.case_prove(gpu_prover, cc, pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)])
    circ = gpu_prover.load_circuit(cc)
    with pytest.raises(_lib.AssertFailed):
        gpu_prover.calculate_witness(circ, [{"out": "5", "bound": "17", "x": "3", "y": "5"}])
    circ.close()


def test_witness_all_circuits(gpu_prover):
    clients = I.simulation_clients(3)
    tr = [c.training_input([0] * 4) for c in clients]
    cases = {
        "balance_unified": [c.balance_input() for c in clients] + [I.balance_integration_input()],
        "sgd_verified": tr + I.sgd_verified_batch(5, nonzero_weights=True),
        "secure_masked_update": [c.secagg_input([j for j in (1, 2, 3) if j != c.id]) for c in clients],
        "secure_agg_client": [I.secure_agg_client_input()],
        "sgd_step_quick": [{k: v for k, v in t.items() if k not in ("weights", "expectedSummedGrad", "remainder", "root_W")}
                           for t in tr],
    }
    for name, ins in cases.items():
        pc.case_witness(gpu_prover, build_circuit(name), ins)


def test_latency_variants_match_the_batch_kernels(gpu_prover, monkeypatch):
    """few instances use one warp per (op, instance) in the witness evaluator and the bit-decomposed bucket reduction; both are
    forced off and on here: same witnesses (vs the oracle) and same proofs (vs the oracle) either way."""
    cc = build_circuit("sgd_verified")
    ins = I.sgd_verified_batch(2) + I.sgd_verified_batch(1, nonzero_weights=True)
    for coop, deep in (("0", "0"), ("1", "1"), ("1", "0")):
        monkeypatch.setenv("ZKFL_WITNESS_COOP", coop)
        monkeypatch.setenv("ZKFL_REDUCE_DEEP", deep)
        pc.case_witness(gpu_prover, cc, ins)
        pc.case_witness(gpu_prover, build_circuit("secure_agg_client"), [I.secure_agg_client_input()])   # Poseidon t = 2, 3, 9
        pc.case_prove(gpu_prover, cc, ins, [(1, 2), (3, 4), (5, 6)], python_verify=0)
        pc.case_g1_msm(gpu_prover, 1 << 14)
        pc.case_g2_msm(gpu_prover, 300)


def test_prove_sgd_verified_batch(gpu_prover):
    """the metric circuit: 8 clients in one batch, fixed (r, s), bit-exact proofs, pairing-verified."""
    ins = I.sgd_verified_batch(6) + I.sgd_verified_batch(2, nonzero_weights=True)
    rnd = random.Random(9)
    rs = [(1, 2)] + [(rnd.randrange(bn.R), rnd.randrange(bn.R)) for _ in range(7)]
    pc.case_prove(gpu_prover, build_circuit("sgd_verified"), ins, rs, python_verify=2)


def test_g2_operand_file_accumulation(gpu_prover, monkeypatch):
    """opt-in operand-file form of the G2 accumulation (shared-memory file, generic operation): bit-exact on the GPU"""
    monkeypatch.setenv("ZKFL_G2_OPERAND_FILE", "1")
    pc.case_g2_msm(gpu_prover, 300)
    pc.case_prove(gpu_prover, build_circuit("sgd_verified"), I.sgd_verified_batch(3, nonzero_weights=True), [(1, 2), (3, 4), (5, 6)])


def test_batch_affine_accumulation_forced(gpu_prover, monkeypatch):
    """The large-batch bucket accumulation (batch-affine, warp-shared inversion) forced on at test sizes: degenerate buckets,
    G1/G2 MSMs against the oracle, and sgd_verified proofs bit-exact with fixed (r, s) for several chunk/slot shapes."""
    monkeypatch.setenv("ZKFL_MSM_AFFINE", "1")
    cc = build_circuit("sgd_verified")
    ins = I.sgd_verified_batch(4) + I.sgd_verified_batch(1, nonzero_weights=True)
    rnd = random.Random(10)
    rs = [(rnd.randrange(bn.R), rnd.randrange(bn.R)) for _ in range(5)]
    for chunk, slots in (("32", "64"), ("8", "5"), ("16", "32")):
        monkeypatch.setenv("ZKFL_MSM_CHUNK", chunk)
        monkeypatch.setenv("ZKFL_MSM_AFFINE_K", slots)
        pc.case_g1_msm_degenerate(gpu_prover)
        for n in (1, 33, 1000, 1 << 14):
            pc.case_g1_msm(gpu_prover, n)
        pc.case_g2_msm(gpu_prover, 300)
        pc.case_prove(gpu_prover, cc, ins, rs, python_verify=1 if slots == "64" else 0)


def test_large_batch_batch_affine_matches_chunk_kernel(gpu_prover, monkeypatch):
    """B = 1024 sgd_verified proofs (the bench shape) through the opt-in batch-affine kernel: its proofs equal the default
    XYZZ chunk kernel's byte for byte, and a sample equals the oracle's."""
    monkeypatch.setenv("ZKFL_MSM_AFFINE", "1")
    cc = build_circuit("sgd_verified")
    circ = gpu_prover.load_circuit(cc)
    zk = gpu_prover.new_zkey(cc, b"affine-1024")
    Z = gpu_prover.load_zkey(zk)
    base = I.sgd_verified_batch(7) + I.sgd_verified_batch(1, nonzero_weights=True)
    B = 1024
    ins = [base[i % 8] for i in range(B)]
    rs = [(i + 1, 7 * i + 3) for i in range(B)]
    p_aff, pub_aff = gpu_prover.full_prove(circ, Z, ins, rs)
    monkeypatch.setenv("ZKFL_MSM_AFFINE", "0")
    p_xyzz, pub_xyzz = gpu_prover.full_prove(circ, Z, ins, rs)
    assert p_aff == p_xyzz and pub_aff == pub_xyzz
    ws = gpu_prover.calculate_witness(circ, base)
    for b in (0, 511, 1023):
        assert (p_aff[b], pub_aff[b]) == ol.groth16_prove(zk, ws[b % 8], *rs[b])
    Z.close()
    circ.close()


def test_prove_other_circuits(gpu_prover):
    clients = I.simulation_clients(3)
    for c in clients:
        c.training_input([0] * 4)
    pc.case_prove(gpu_prover, build_circuit("balance_unified"), [c.balance_input() for c in clients[:2]], [(7, 8), (9, 10)])
    pc.case_prove(gpu_prover, build_circuit("secure_masked_update"),
                  [c.secagg_input([j for j in (1, 2, 3) if j != c.id]) for c in clients], [(1, 1), (2, 3), (5, 8)])
    pc.case_prove(gpu_prover, build_circuit("secure_agg_client"), [I.secure_agg_client_input()], [(4, 4)])


def test_random_blinding_still_verifies(gpu_prover):
    """rs = NULL -> r, s from the OS like snarkjs; proofs differ per call but verify."""
    import groth16_ref as g16
    from zkfl_b200.formats import export_verification_key
    cc = pc.tiny_circuit()
    circ = gpu_prover.load_circuit(cc)
    zk = gpu_prover.new_zkey(cc.r1cs_bytes(), b"rnd")
    Z = gpu_prover.load_zkey(zk)
    p1, pub1 = gpu_prover.full_prove(circ, Z, pc.tiny_inputs()[:1])
    p2, _ = gpu_prover.full_prove(circ, Z, pc.tiny_inputs()[:1])
    assert p1 != p2
    vk = g16.vkey_from_json(export_verification_key(zk))
    assert g16.verify(vk, ol.ints(pub1[0]), g16.proof_from_bytes(p1[0]))
    assert g16.verify(vk, ol.ints(pub1[0]), g16.proof_from_bytes(p2[0]))


def test_product_verifier_agrees_with_oracle(gpu_prover):
    import groth16_ref as g16
    from zkfl_b200 import formats
    from zkfl_b200 import snarkjs as sj
    cc = build_circuit("secure_masked_update")
    clients = I.simulation_clients(3)
    for c in clients:
        c.training_input([0] * 4)
    ins = [c.secagg_input([j for j in (1, 2, 3) if j != c.id]) for c in clients]
    zk, proofs, pubs = pc.case_prove(gpu_prover, cc, ins, [(1, 2), (3, 4), (5, 6)], python_verify=1)
    vkj = formats.export_verification_key(zk)
    for p, q in zip(proofs, pubs):
        sig = formats.publics_bytes_to_json(q)
        assert sj.groth16.verify(vkj, sig, formats.proof_bytes_to_json(p))
        assert not sj.groth16.verify(vkj, [str(int(sig[0]) + 1)] + sig[1:], formats.proof_bytes_to_json(p))
    swapped = formats.proof_bytes_to_json(proofs[0])
    assert not sj.groth16.verify(vkj, formats.publics_bytes_to_json(pubs[1]), swapped)


def test_batch_verifier_on_gpu(gpu_prover):
    """all proofs of a round verified in one GPU pass.  Valid and tampered sgd_verified proofs against the
    oracle's pairing check and the single-proof host verifier; then 1024 and 3072 proofs (a 1024-client round has 3072) with
    every 7th one corrupted, timed."""
    import time
    from zkfl_b200 import formats
    cc = build_circuit("sgd_verified")
    ins = I.sgd_verified_batch(3) + I.sgd_verified_batch(1, nonzero_weights=True)
    zk, proofs, pubs = pc.case_prove(gpu_prover, cc, ins, [(1, 2), (3, 4), (5, 6), (7, 8)], python_verify=0)
    pc.case_verify_batch(gpu_prover, zk, proofs, pubs)
    vk = formats.vkey_json_to_bytes(formats.export_verification_key(zk))
    for B in (1024, 3072):
        ps = [proofs[i % 4] for i in range(B)]
        qs = [pubs[i % 4] for i in range(B)]
        for i in range(0, B, 7):
            ps[i] = ps[i][:192] + ps[(i + 1) % B][192:] if i % 14 else ps[i][:64] + proofs[(i + 1) % 4][64:192] + ps[i][192:]
        expect = [(i % 7 != 0) or ps[i] == proofs[i % 4] for i in range(B)]
        gpu_prover.verify_batch(vk, qs[:8], ps[:8])      # workspace
        gpu_prover.prof_enable(True)
        t = time.time()
        got = gpu_prover.verify_batch(vk, qs, ps)
        dt = time.time() - t
        prof = gpu_prover.prof_read()
        gpu_prover.prof_enable(False)
        assert got == expect
        print(f"verify_batch B={B}: {dt * 1e3:.1f} ms ({B / dt:.0f} proofs/s)", {k: round(v["ms"], 1) for k, v in prof.items() if k.startswith("verify")})


def test_split_msm_partials_single_gpu(gpu_prover):
    """the multi-GPU split of one proof (point ranges + gather + add), with the ranks emulated one after another on one GPU"""
    cc = build_circuit("sgd_verified")
    circ = gpu_prover.load_circuit(cc)
    zk = gpu_prover.new_zkey(cc, b"split")
    Z = gpu_prover.load_zkey(zk)
    ins = I.sgd_verified_batch(2, nonzero_weights=True)
    ws = gpu_prover.calculate_witness(circ, ins)
    rs = [(123, 456), (789, 1011)]
    for nparts in (2, 3, 8):
        parts = [gpu_prover.msm_partials(Z, ws, r, nparts) for r in range(nparts)]
        got = gpu_prover.finalize(Z, parts, 2, rs)
        assert got == [ol.groth16_prove(zk, w, *r)[0] for w, r in zip(ws, rs)], nparts
    Z.close()
    Z = gpu_prover.load_zkey(zk, nparts=8)        # zkfl_zkey_load_split: smaller windows, same proof bytes
    parts = [gpu_prover.msm_partials(Z, ws, r, 8) for r in range(8)]
    assert gpu_prover.finalize(Z, parts, 2, rs) == got
    Z.close()
    circ.close()


def test_balance_unified_prod_2pow19(gpu_prover):
    """BASELINE configs[0] at production size: BalanceProofUnified(128,7,16) over the seed-42 dataset whose Merkle root is
    data/test_input_v5.json's root_D; 362k wires, domain 2^19; bit-exact against the oracle and pairing-verified."""
    import json
    import os
    import groth16_ref as g16
    from zkfl_b200 import formats
    from zkfl_b200 import snarkjs as sj
    inp = I.balance_prod_input()
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "test_input_v5.json")))
    assert inp["root"] == golden["root_D"]
    cc = build_circuit("balance_unified_prod")
    circ = gpu_prover.load_circuit(cc, check_constraints=False)
    zk = gpu_prover.new_zkey(cc, b"prod")
    Z = gpu_prover.load_zkey(zk)
    assert Z.domain == 1 << 19
    ws = gpu_prover.calculate_witness(circ, [inp])
    assert ws[0] == ol.witness_batch(cc.program_bytes(), circ.pack_inputs([inp]), cc.n_inputs, cc.n_wires)
    proofs, pubs = gpu_prover.prove(Z, ws, [(7, 9)])
    ref_p, ref_pub = ol.groth16_prove(zk, ws[0], 7, 9)
    assert proofs[0] == ref_p and pubs[0] == ref_pub
    sig = formats.publics_bytes_to_json(pubs[0])
    assert sig[1] == golden["root_D"] and sig[2] == "128"
    assert sj.groth16.verify(formats.export_verification_key(zk), sig, formats.proof_bytes_to_json(proofs[0]))
    Z.close()
    circ.close()


def test_full_round_like_the_reference_simulation(gpu_prover):
    """tests/full_system_simulation.mjs flow (3 clients, 9 proofs, 9 verifications, masked aggregation) on the GPU backend,
    then a 24-client round (72 proofs, BASELINE configs[3] shape) without re-running setup."""
    from zkfl_b200 import simulation
    cache = {}
    rep = simulation.run_round(gpu_prover, 3, cache=cache)
    assert rep["verified"] == {"balance": 3, "training": 3, "secagg": 3}
    assert rep["aggregated_gradient"] == rep["expected_gradient"]           # the pairwise masks cancel
    assert rep["new_model"] == [-0.01 * g for g in rep["aggregated_gradient"]]
    rep = simulation.run_round(gpu_prover, 24, cache=cache)
    assert rep["verified"] == {"balance": 24, "training": 24, "secagg": 24}
    assert rep["aggregated_gradient"] == rep["expected_gradient"]
    host = simulation.run_round(gpu_prover, 24, cache=cache, gpu_inputs=False)      # per-client host Poseidon: same round
    assert host["verified"] == rep["verified"] and host["aggregated_gradient"] == rep["aggregated_gradient"]
    assert host["new_model"] == rep["new_model"]


def test_masked_aggregation_and_model_update_on_gpu(gpu_prover):
    from test_emul_pipeline import case_aggregate
    case_aggregate(gpu_prover)


def _host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().total / 2 ** 30
    except Exception:
        return 0


@pytest.mark.skipif(_host_ram_gb() < 48 or bool(__import__("os").environ.get("ZKFL_SKIP_SLOW")),
                    reason="needs ~20 GB of host RAM for the 1.6 GB key (about one minute on a B200 box)")
def test_scaled_training_circuit_2pow20_single_proof_and_split(gpu_prover):
    """BASELINE configs[4]: TrainingStepVerified(256, 32, 8, 1000) -- 976k wires, 966k constraints, domain 2^20 -- as ONE
    proof: whole on one context, and with every MSM split into 8 point ranges (the 8-GPU exchange, ranks emulated one after
    another on this GPU); both bit-exact against the oracle and verified."""
    import time
    from zkfl_b200 import formats
    from zkfl_b200 import snarkjs as sj
    from zkfl_b200.circuits.library import training_step_verified
    t = time.time()
    cc = training_step_verified(256, 32, 8, 1000, "sgd_scaled_2pow20")
    inp = I.scaled_training_input(256, 32, 8)
    circ = gpu_prover.load_circuit(cc, check_constraints=False)
    zk = gpu_prover.new_zkey(cc, b"scaled")
    Z = gpu_prover.load_zkey(zk)
    assert Z.domain == 1 << 20
    print(f"build + setup {time.time() - t:.0f} s, zkey {len(zk) / 1e6:.0f} MB")
    t = time.time()
    ws = gpu_prover.calculate_witness(circ, [inp], check=False)
    proofs, pubs = gpu_prover.prove(Z, ws, [(3, 4)])
    print(f"witness + prove (first call, buffers allocated) {time.time() - t:.2f} s")
    t = time.time()
    proofs2, _ = gpu_prover.prove(Z, ws, [(3, 4)])
    print(f"prove, second call {time.time() - t:.3f} s")
    assert proofs2 == proofs
    t = time.time()
    ref_p, ref_pub = ol.groth16_prove(zk, ws[0], 3, 4)
    print(f"oracle prove {time.time() - t:.1f} s on {ol.ncores()} threads")
    assert ws[0] == ol.witness_batch(cc.program_bytes(), circ.pack_inputs([inp]), cc.n_inputs, cc.n_wires)
    assert proofs[0] == ref_p and pubs[0] == ref_pub
    parts = [gpu_prover.msm_partials(Z, ws, r, 8) for r in range(8)]
    assert gpu_prover.finalize(Z, parts, 1, [(3, 4)]) == [ref_p]
    Z.close()
    Z = gpu_prover.load_zkey(zk, nparts=8)        # the key as the 8 ranks load it (zkfl_zkey_load_split)
    t = time.time()
    parts = [gpu_prover.msm_partials(Z, ws, r, 8) for r in range(8)]
    assert gpu_prover.finalize(Z, parts, 1, [(3, 4)]) == [ref_p]
    print(f"8 emulated ranks with the split key: {time.time() - t:.3f} s")
    sig = formats.publics_bytes_to_json(pubs[0])
    assert sj.groth16.verify(formats.export_verification_key(zk), sig, formats.proof_bytes_to_json(proofs[0]))
    Z.close()
    circ.close()


def test_commitment_pipeline_on_gpu_for_1024_clients(gpu_prover):
    """dataset Merkle trees, commitments and PRF masks of 1024 clients in one batched run; a sample of the
    clients is compared with the oracle's restatement of the reference's JavaScript helpers."""
    import time
    from zkfl_b200 import commitments
    lcg = I.JsLcg(777)
    req = []
    for cid in range(1, 1025):
        feats = [[lcg.random_int(0, 100) for _ in range(4)] for _ in range(8)]
        labels = [(i + cid) % 2 for i in range(8)]
        base = 3 * ((cid - 1) // 3)
        peers = [base + k for k in (1, 2, 3) if base + k != cid]
        req.append({"features": feats, "labels": labels, "weights": [lcg.random_int(-1000, 999) for _ in range(4)],
                    "gradient": [lcg.random_int(-50, 49) for _ in range(4)], "client_id": cid, "round": 1,
                    "master_key": 1000 + 111 * cid, "peer_ids": peers, "shared_keys": [cid * 7 + j for j in peers]})
    t = time.time()
    got = commitments.compute(gpu_prover, req, 8, 4, 3)
    print(f"commitments of 1024 clients: {time.time() - t:.2f} s (host packing included)")
    for idx in (0, 1, 511, 1023):
        r, g = req[idx], got[idx]
        tree = bn.build_merkle_tree([bn.vector_hash(f + [l]) for f, l in zip(r["features"], r["labels"])], 3)
        assert g["tree"] == tree
        assert g["root_W"] == bn.weight_commitment(r["weights"])
        assert g["root_G"] == bn.gradient_commitment([x % bn.R for x in r["gradient"]], r["client_id"], 1)
        assert g["root_K"] == bn.key_material_commitment(r["master_key"], r["shared_keys"])
        assert g["masks"][1] == bn.derive_pairwise_mask(r["shared_keys"][1], 1, r["client_id"], r["peer_ids"][1], 4)


def test_odd_batch_sizes_partial_warps(gpu_prover):
    """B = 1, 33, 65: batch-minor layout with partially filled warps; every proof against the oracle"""
    cc = build_circuit("sgd_step_quick")
    circ = gpu_prover.load_circuit(cc)
    zk = gpu_prover.new_zkey(cc, b"odd")
    Z = gpu_prover.load_zkey(zk)
    clients = I.simulation_clients(3)
    base = []
    for c in clients:
        t = c.training_input([0] * 4)
        base.append({k: v for k, v in t.items() if k not in ("weights", "expectedSummedGrad", "remainder", "root_W")})
    rnd = random.Random(33)
    for B in (1, 33, 65):
        ins = [base[i % 3] for i in range(B)]
        rs = [(rnd.randrange(bn.R), rnd.randrange(bn.R)) for _ in range(B)]
        proofs, pubs = gpu_prover.full_prove(circ, Z, ins, rs)
        ws = gpu_prover.calculate_witness(circ, base)
        for b in (0, B // 2, B - 1):
            ref_p, ref_pub = ol.groth16_prove(zk, ws[b % 3], *rs[b])
            assert proofs[b] == ref_p and pubs[b] == ref_pub, (B, b)
    Z.close()
    circ.close()


def test_secure_aggregation_batch_of_256(gpu_prover):
    """BASELINE configs[2]: 256 SecureMaskedUpdate(4,2) client proofs in one batch (federations of three, seeded gradients in
    [-50, 49] as tests/test_secure_aggregation.mjs:151); all verified, a sample bit-exact against the oracle, masks cancel."""
    import time
    from zkfl_b200 import commitments, formats
    from zkfl_b200 import snarkjs as sj
    n = 255   # 85 federations of three
    lcg = I.JsLcg(4242)
    req = []
    for cid in range(1, n + 1):
        base = 3 * ((cid - 1) // 3)
        peers = [base + k for k in (1, 2, 3) if base + k != cid]
        req.append({"features": [[0] * 4] * 8, "labels": [0] * 8, "weights": [0] * 4, "gradient": [lcg.random_int(-50, 49) for _ in range(4)],
                    "client_id": cid, "round": 1, "master_key": 1000 + 111 * cid, "peer_ids": peers,
                    "shared_keys": [bn.poseidon([min(cid, j), max(cid, j), 12345]) for j in peers]})
    com = commitments.compute(gpu_prover, req, 8, 4, 3)      # root_G, root_K and the PRF masks of all clients on the GPU
    root_d, root_w = bn.poseidon([12345]), bn.poseidon([67890])  # test_secure_aggregation.mjs:275-276
    ins = []
    for r, c in zip(req, com):
        masked = [g % bn.R for g in r["gradient"]]
        for j, mask in zip(r["peer_ids"], c["masks"]):
            masked = [(m + x) % bn.R if r["client_id"] < j else (m - x) % bn.R for m, x in zip(masked, mask)]
        r["masked"] = masked
        ins.append({"client_id": r["client_id"], "round": 1, "root_D": root_d, "root_G": c["root_G"], "root_W": root_w, "root_K": c["root_K"],
                    "tauSquared": 100000000, "masked_update": masked, "peer_ids": r["peer_ids"], "gradient": r["gradient"],
                    "master_key": r["master_key"], "shared_keys": r["shared_keys"]})
    for f in range(0, n, 3):   # the pairwise masks cancel inside every federation (test_secure_aggregation.mjs:215-238)
        for k in range(4):
            assert sum(req[f + i]["masked"][k] for i in range(3)) % bn.R == sum(req[f + i]["gradient"][k] for i in range(3)) % bn.R
    cc = build_circuit("secure_masked_update")
    circ = gpu_prover.load_circuit(cc)
    zk = gpu_prover.new_zkey(cc, b"secagg256")
    Z = gpu_prover.load_zkey(zk)
    ws = gpu_prover.calculate_witness(circ, ins)            # every === holds for all 255 clients
    rs = [(i + 1, 1000 + i) for i in range(n)]
    t = time.time()
    proofs, pubs = gpu_prover.full_prove(circ, Z, ins, rs)
    print(f"{n} secure_masked_update proofs: {time.time() - t:.3f} s")
    for b in (0, 100, n - 1):
        assert (proofs[b], pubs[b]) == ol.groth16_prove(zk, ws[b], *rs[b])
    vk = formats.export_verification_key(zk)
    ok = sj.groth16.verifyBatch(vk, [(formats.publics_bytes_to_json(q), formats.proof_bytes_to_json(p)) for p, q in zip(proofs, pubs)], prover=gpu_prover)
    assert all(ok) and len(ok) == n
    Z.close()
    circ.close()


# ---------------------------------------------------------------------------------------------------------------------
# the reference's OWN fixture on the CUDA path, and the two parity checks that ran on the CPU only
def _v5():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "test_input_v5.json")))


def test_reference_fixture_sgd_step_v5_on_gpu(gpu_prover):
    """`sgd_step_v5(8,16,7)` (src/circuits/training/sgd_step_v5.circom:86-168) over data/test_input_v5.json -- the only input the
    reference tree holds (BASELINE configs[0] names the file).  On the device: witness == both oracle interpreters (C++ and
    pure Python), public signals == [client_id, round, root_D, root_G, tauSquared] of the fixture, every constraint holds
    (device check AND the Python R1CS checker on the `.r1cs` bytes), proof bit-exact vs the oracle with fixed (r, s) at B = 1 and
    for a tiled batch of 64, accepted by the oracle's pairing, the host verifier and the GPU batch verifier."""
    import groth16_ref as g16
    import witness_ref as wr
    from zkfl_b200 import formats
    from zkfl_b200 import snarkjs as sj
    d = _v5()
    cc = build_circuit("sgd_step_v5")
    circ = gpu_prover.load_circuit(cc)
    ws = gpu_prover.calculate_witness(circ, [d])                     # device witness + device constraint check
    assert ws[0] == ol.witness_batch(cc.program_bytes(), circ.pack_inputs([d]), cc.n_inputs, cc.n_wires)
    w_int = ol.ints(ws[0])
    assert w_int == wr.calculate_witness(wr.Program(cc.program_bytes()), cc.flatten_input(d))
    assert wr.R1cs(cc.r1cs_bytes()).first_violation(w_int) is None
    expect_pub = [1, 1, int(d["root_D"]), int(d["root_G"]), int(d["tauSquared"])]
    assert w_int[1:6] == expect_pub and int(d["tauSquared"]) == 76014
    zk = gpu_prover.new_zkey(cc, b"v5-fixture")
    Z = gpu_prover.load_zkey(zk)
    assert Z.domain == 1 << 15
    ref1 = ol.groth16_prove(zk, ws[0], 11, 13)
    p1, q1 = gpu_prover.full_prove(circ, Z, [d], [(11, 13)])         # B = 1: latency kernels
    assert (p1[0], q1[0]) == ref1 and ol.ints(q1[0]) == expect_pub
    assert gpu_prover.prove(Z, ws, [(11, 13)])[0][0] == ref1[0]
    B = 64                                                            # tiled batch: throughput kernels
    rs = [(11, 13)] + [(1000 + b, 7 * b + 1) for b in range(1, B)]
    pB, qB = gpu_prover.full_prove(circ, Z, [d] * B, rs)
    assert pB[0] == ref1[0] and all(q == ref1[1] for q in qB)
    for b in (1, 31, 63):
        assert pB[b] == ol.groth16_prove(zk, ws[0], *rs[b])[0]
    vkj = formats.export_verification_key(zk)
    assert g16.verify(g16.vkey_from_json(vkj), expect_pub, g16.proof_from_bytes(p1[0]))
    sig = formats.publics_bytes_to_json(q1[0])
    assert sig == [str(v) for v in expect_pub]
    assert sj.groth16.verify(vkj, sig, formats.proof_bytes_to_json(p1[0]))
    assert gpu_prover.verify_batch(formats.vkey_json_to_bytes(vkj), qB, pB) == [True] * B
    # tampered public inputs: the circuit's === fails ON THE DEVICE, in the witness call and inside the one-pass fullProve
    for key, delta in (("root_G", 1), ("root_D", 1), ("tauSquared", -70000)):
        t = dict(d)
        t[key] = str(int(d[key]) + delta)
        with pytest.raises(_lib.AssertFailed):
            gpu_prover.calculate_witness(circ, [t])
        with pytest.raises(_lib.AssertFailed):
            gpu_prover.full_prove(circ, Z, [d, t], [(1, 2), (3, 4)])
    Z.close()
    circ.close()


def test_product_setup_matches_the_oracle_setup_on_gpu(gpu_prover):
    """`groth16 setup` replacement on the REAL library (device scalar multiplications) against the oracle's own setup from the
    same toxic waste: header and every point section byte-identical, coefficient sections equal as sets -- for the fixture's
    circuit and the metric circuit."""
    import struct
    import groth16_ref as g16
    import witness_ref as wr
    from zkfl_b200.zkey_setup import toxic_from_seed

    def coeff_set(sec):
        n = struct.unpack_from("<I", sec, 0)[0]
        return {sec[4 + 44 * i:48 + 44 * i] for i in range(n)}
    for name in ("sgd_step_v5", "sgd_verified"):
        cc = build_circuit(name)
        mine = gpu_prover.new_zkey(cc, b"gpu-setup-parity")
        ref = g16.setup_fast(wr.R1cs(cc.r1cs_bytes()), *toxic_from_seed(b"gpu-setup-parity"))
        _, a = wr.read_sections(mine, b"zkey")
        _, b = wr.read_sections(ref, b"zkey")
        for sid in (1, 2, 3, 5, 6, 7, 8, 9):
            assert a[sid] == b[sid], f"{name}: section {sid} differs"
        assert coeff_set(a[4]) == coeff_set(b[4]) and len(a[4]) == len(b[4])


def test_full_prove_checks_constraints_in_one_pass(gpu_prover):
    """full_prove validates inputs and runs the R1CS check on the HBM-resident witness; a failing batch
    returns ZKFL_ERR_ASSERT with first_bad per instance and no proof bytes; malformed witnesses are refused by prove."""
    import ctypes
    cc = build_circuit("sgd_verified")
    circ = gpu_prover.load_circuit(cc)
    Z = gpu_prover.load_zkey(gpu_prover.new_zkey(cc, b"check"))
    good = I.sgd_verified_batch(3)
    bad = dict(good[1])
    bad["gradPos"] = ["1"] + good[1]["gradPos"][1:]
    packed = circ.pack_inputs([good[0], bad, good[2]])
    first_bad = (ctypes.c_uint32 * 3)()
    out_p = ctypes.create_string_buffer(b"\x01" * 768, 768)
    before = gpu_prover.launch_count()
    rc = gpu_prover.lib.zkfl_groth16_full_prove_batch(gpu_prover.ctx, circ.handle, Z.handle, circ.r1cs_handle, _lib.as_ptr(packed), None, 3,
                                                      out_p, None, first_bad)
    assert rc == -5 and out_p.raw == bytes(768)
    assert first_bad[0] == 0xFFFFFFFF and first_bad[1] != 0xFFFFFFFF and first_bad[2] == 0xFFFFFFFF
    assert gpu_prover.launch_count() > before
    proofs, _ = gpu_prover.full_prove(circ, Z, good, [(1, 2), (3, 4), (5, 6)])
    ws = gpu_prover.calculate_witness(circ, good)
    assert proofs == gpu_prover.prove(Z, ws, [(1, 2), (3, 4), (5, 6)])[0]
    with pytest.raises(_lib.ZkflError, match="wire 0"):
        gpu_prover.prove(Z, [bytes(32) + ws[0][32:]], [(1, 2)])
    with pytest.raises(_lib.ZkflError, match="not reduced"):
        gpu_prover.prove(Z, [ws[0][:32 * 50] + b"\xff" * 32 + ws[0][32 * 51:]], [(1, 2)])
    with pytest.raises(_lib.ZkflError, match="not reduced"):
        gpu_prover.full_prove(circ, Z, b"\xff" * (32 * circ.n_inputs), [(1, 2)])
    Z.close()
    circ.close()


def test_contributed_key_on_gpu(gpu_prover):
    """`zkey contribute` on the device (point sections rescaled by the fresh secret): proofs under the contributed key are
    bit-exact against the oracle reading the same key bytes and verify under its verification key only."""
    import groth16_ref as g16
    from zkfl_b200 import formats
    cc = build_circuit("secure_agg_client")
    base = gpu_prover.new_zkey(cc)                                    # OS randomness: not reproducible
    assert base != gpu_prover.new_zkey(cc)
    final = gpu_prover.contribute_zkey(base, "contributor", b"entropy")
    circ = gpu_prover.load_circuit(cc)
    Z = gpu_prover.load_zkey(final)
    ins = [I.secure_agg_client_input()]
    proofs, pubs = gpu_prover.full_prove(circ, Z, ins, [(21, 34)])
    w = gpu_prover.calculate_witness(circ, ins)[0]
    assert (proofs[0], pubs[0]) == ol.groth16_prove(final, w, 21, 34)
    vk_new = formats.vkey_json_to_bytes(formats.export_verification_key(final))
    vk_old = formats.vkey_json_to_bytes(formats.export_verification_key(base))
    assert gpu_prover.verify_batch(vk_new, pubs, proofs) == [True]
    assert gpu_prover.verify_batch(vk_old, pubs, proofs) == [False]
    assert g16.verify(g16.vkey_from_json(formats.export_verification_key(final)), ol.ints(pubs[0]), g16.proof_from_bytes(proofs[0]))
    Z.close()
    circ.close()
