"""CPU suite for the kernel logic: the CUDA sources compiled as a host emulation (tests/_emul, a test
double that the product never loads) must agree with the oracle bit for bit. The GPU suite
(tests/test_gpu_parity.py) repeats these cases on the real sm_100a library at full sizes."""
import os

import pytest

import parity_cases as pc
from zkfl_b200 import _lib
from zkfl_b200 import inputs as I
from zkfl_b200.circuits import build_circuit


def test_generator_mul(emul_prover):
    pc.case_generator_mul(emul_prover, n1=6, n2=2)


def test_g1_msm_sizes_and_edges(emul_prover):
    for n in (1, 2, 33, 300):
        pc.case_g1_msm(emul_prover, n)
    pc.case_g1_msm_degenerate(emul_prover)
    pc.case_linearity(emul_prover, n=32)


def test_g2_msm(emul_prover):
    pc.case_g2_msm(emul_prover, 40)


def test_both_bucket_reductions(emul_prover, monkeypatch):
    """few rows take the bit-decomposed tree of plain sums (latency variant), batches the three-level running sums; both forced
    here on the same MSMs and proofs"""
    for mode in ("0", "1"):
        monkeypatch.setenv("ZKFL_REDUCE_DEEP", mode)
        for n in (1, 33, 300):
            pc.case_g1_msm(emul_prover, n)
        pc.case_g1_msm_degenerate(emul_prover)
        pc.case_g2_msm(emul_prover, 40)
        pc.case_prove(emul_prover, pc.tiny_circuit(), pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)], python_verify=0)


def test_tiny_circuit_witness_prove_verify(emul_prover):
    cc = pc.tiny_circuit()
    pc.case_witness(emul_prover, cc, pc.tiny_inputs())
    pc.case_prove(emul_prover, cc, pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)])


def test_resident_msm_with_window_table(emul_prover):
    pc.case_msm_resident(emul_prover, 1030, 1)
    pc.case_msm_resident(emul_prover, 1024, 2)
    pc.case_msm_resident(emul_prover, 40, 1)        # below the table threshold


def test_queued_fixup_of_large_batches(emul_prover, monkeypatch):
    """The fix-up of large batches (cut buckets queued by cut count, summed by k_msm_fixup_apply) forced on small cases with short
    chunks, so that buckets are whole, cut once and cut several times: MSMs with degenerate buckets and a proof batch, bit-exact."""
    monkeypatch.setenv("ZKFL_FIXUP_QUEUE_MIN_ROWS", "1")
    for chunk in ("4", "7", "64"):
        monkeypatch.setenv("ZKFL_MSM_CHUNK", chunk)
        pc.case_g1_msm_degenerate(emul_prover)
        for n in (1, 33, 300):
            pc.case_g1_msm(emul_prover, n)
        pc.case_g2_msm(emul_prover, 40)
        cc = pc.tiny_circuit()
        pc.case_prove(emul_prover, cc, pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)])


def test_g2_operand_file_accumulation(emul_prover, monkeypatch):
    """The opt-in operand-file form of the G2 bucket accumulation (running sum and temporaries in a per-thread file, one generic
    operation as the only call): G2 MSMs and a proof batch, bit-exact against the oracle, with short and default chunks."""
    monkeypatch.setenv("ZKFL_G2_OPERAND_FILE", "1")
    for chunk in ("4", "64"):
        monkeypatch.setenv("ZKFL_MSM_CHUNK", chunk)
        pc.case_g2_msm(emul_prover, 40)
        pc.case_prove(emul_prover, pc.tiny_circuit(), pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)])


def test_batch_affine_accumulation(emul_prover, monkeypatch):
    """The batch-affine bucket accumulation (large-batch path) forced on small cases: degenerate buckets (doubling,
    P + (-P), infinity bases, runs cut by chunk borders) and a whole proof batch, bit-exact against the oracle."""
    monkeypatch.setenv("ZKFL_MSM_AFFINE", "1")
    for chunk, slots in (("4", "3"), ("8", "64")):
        monkeypatch.setenv("ZKFL_MSM_CHUNK", chunk)
        monkeypatch.setenv("ZKFL_MSM_AFFINE_K", slots)
        pc.case_g1_msm_degenerate(emul_prover)
        for n in (1, 33, 300):
            pc.case_g1_msm(emul_prover, n)
        pc.case_g2_msm(emul_prover, 40)
    cc = pc.tiny_circuit()
    pc.case_prove(emul_prover, cc, pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)])


def test_batch_verifier_matches_oracle_and_host_verifier(emul_prover, monkeypatch):
    import __graft_entry__ as ge
    from zkfl_b200 import snarkjs as sj
    monkeypatch.setitem(sj._state, "lib_path", ge.build_emul())   # the single-proof verifier of the same (emulated) build
    assert emul_prover.lib.zkfl_debug_pairing_selftest() == 0     # tower view of Fq12 against the flat basis
    cc = pc.tiny_circuit()
    zk, proofs, pubs = pc.case_prove(emul_prover, cc, pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)], python_verify=0)
    pc.case_verify_batch(emul_prover, zk, proofs, pubs)            # default: easy part + x-power chain
    monkeypatch.setenv("ZKFL_VERIFY_FLAT", "1")                     # the inversion-free two-power form agrees
    pc.case_verify_batch(emul_prover, zk, proofs, pubs)


def test_single_proof_entry_points_and_json_writers(emul_prover):
    """SURVEY 8b's single-proof forms (zkfl_wtns_calculate / zkfl_groth16_prove / zkfl_groth16_full_prove) equal the batch
    calls with B = 1, and zkfl_proof_to_json / zkfl_public_to_json write snarkjs's proof.json / public.json."""
    import ctypes
    import json
    from zkfl_b200 import formats
    P, lib = emul_prover, emul_prover.lib
    cc = pc.tiny_circuit()
    circ = P.load_circuit(cc)
    zk = P.new_zkey(cc, b"single")
    Z = P.load_zkey(zk)
    packed = circ.pack_inputs(pc.tiny_inputs()[:1])
    ws = P.calculate_witness(circ, packed)
    proofs, pubs = P.prove(Z, ws, [(5, 9)])
    w1 = ctypes.create_string_buffer(32 * cc.n_wires)
    bad = (ctypes.c_uint32 * 1)()
    P._check(lib.zkfl_wtns_calculate(P.ctx, circ.handle, circ.r1cs_handle, _lib.as_ptr(packed), w1, bad))
    assert w1.raw == ws[0]
    r, s_ = (5).to_bytes(32, "little"), (9).to_bytes(32, "little")
    p1, q1 = ctypes.create_string_buffer(256), ctypes.create_string_buffer(32 * Z.n_public)
    P._check(lib.zkfl_groth16_prove(P.ctx, Z.handle, _lib.as_ptr(ws[0]), _lib.as_ptr(r), _lib.as_ptr(s_), p1, q1))
    assert (p1.raw, q1.raw) == (proofs[0], pubs[0])
    p2, q2 = ctypes.create_string_buffer(256), ctypes.create_string_buffer(32 * Z.n_public)
    P._check(lib.zkfl_groth16_full_prove(P.ctx, circ.handle, Z.handle, circ.r1cs_handle, _lib.as_ptr(packed), _lib.as_ptr(r), _lib.as_ptr(s_), p2, q2))
    assert (p2.raw, q2.raw) == (proofs[0], pubs[0])
    p3 = ctypes.create_string_buffer(256)                                   # r, s = NULL: random blinding, still a valid proof
    P._check(lib.zkfl_groth16_prove(P.ctx, Z.handle, _lib.as_ptr(ws[0]), None, None, p3, q2))
    assert p3.raw != proofs[0] and P.verify_batch(formats.vkey_json_to_bytes(formats.export_verification_key(zk)), [pubs[0]], [p3.raw]) == [True]
    # one contiguous witness buffer (bytearray / torch tensor: pinned or device memory on the GPU) instead of a list
    import torch
    two = P.calculate_witness(circ, circ.pack_inputs(pc.tiny_inputs()[:2]))
    ref2 = P.prove(Z, two, [(5, 9), (6, 7)])
    assert P.prove(Z, bytearray(b"".join(two)), [(5, 9), (6, 7)]) == ref2
    assert P.prove(Z, torch.frombuffer(bytearray(b"".join(two)), dtype=torch.uint8), [(5, 9), (6, 7)]) == ref2
    with pytest.raises(ValueError):
        P.prove(Z, bytearray(b"".join(two))[:-1], [(5, 9), (6, 7)])
    buf = ctypes.create_string_buffer(4096)
    P._check(lib.zkfl_proof_to_json(_lib.as_ptr(proofs[0]), buf, len(buf)))
    assert json.loads(buf.value) == formats.proof_bytes_to_json(proofs[0])
    P._check(lib.zkfl_public_to_json(_lib.as_ptr(pubs[0]), Z.n_public, buf, len(buf)))
    assert json.loads(buf.value) == formats.publics_bytes_to_json(pubs[0])
    edge = b"".join(v.to_bytes(32, "little") for v in (0, 1, 10 ** 9, 10 ** 18 - 1, 2 ** 256 - 1))
    P._check(lib.zkfl_public_to_json(_lib.as_ptr(edge), 5, buf, len(buf)))
    assert json.loads(buf.value) == [str(v) for v in (0, 1, 10 ** 9, 10 ** 18 - 1, 2 ** 256 - 1)]
    assert lib.zkfl_proof_to_json(_lib.as_ptr(proofs[0]), buf, 10) != 0     # buffer too small: an error, not a truncation
    Z.close()
    circ.close()


def test_failed_constraint_raises_assert(emul_prover):
    cc = pc.tiny_circuit()
    circ = emul_prover.load_circuit(cc)
    with pytest.raises(_lib.AssertFailed):
        emul_prover.calculate_witness(circ, [{"out": "5", "bound": "17", "x": "3", "y": "5"}])
    with pytest.raises(_lib.AssertFailed):   # x >= bound
        emul_prover.calculate_witness(circ, [{"out": str(15 ** 2 + 3), "bound": "3", "x": "3", "y": "5"}])
    circ.close()


def test_sgd_verified_witness_batch(emul_prover):
    cc = build_circuit("sgd_verified")
    pc.case_witness(emul_prover, cc, I.sgd_verified_batch(2) + I.sgd_verified_batch(1, nonzero_weights=True))


def test_bad_arguments_fail_loudly(emul_prover):
    with pytest.raises(_lib.ZkflError):
        emul_prover.load_zkey(b"zkey" + bytes(100))
    with pytest.raises(_lib.ZkflError):
        emul_prover.g1_mul_generator((2 ** 256 - 1).to_bytes(32, "little"))   # not reduced mod r


def test_snarkjs_surface_and_cli_round_trip(tmp_path, monkeypatch):
    """The file-based flow of the reference's tests (compile, setup, witness, prove, verify) through the snarkjs-named
    API and the CLI shim, on the emulation library; the proof must match the oracle and verify in both verifiers."""
    import json
    import __graft_entry__ as ge
    import groth16_ref as g16
    import oracle_lib as ol
    from zkfl_b200 import snarkjs as sj
    from zkfl_b200 import cli, formats
    monkeypatch.setitem(sj._state, "lib_path", ge.build_emul())   # test double, handed over in code (no environment redirect exists)
    monkeypatch.setitem(sj._state, "prover", None)
    monkeypatch.setitem(sj._state, "circuits", {})
    monkeypatch.setitem(sj._state, "zkeys", {})
    d = str(tmp_path)
    assert cli.main(["circom", "secure_agg_client.circom", "--r1cs", "--wasm", "--sym", "-o", d]) == 0
    assert cli.main(["snarkjs", "r1cs", "info", f"{d}/secure_agg_client.r1cs"]) == 0
    assert cli.main(["snarkjs", "groth16", "setup", f"{d}/secure_agg_client.r1cs", f"{d}/pot12_final.ptau", f"{d}/c_0000.zkey"]) == 0
    assert cli.main(["snarkjs", "zkey", "contribute", f"{d}/c_0000.zkey", f"{d}/c_final.zkey", "--name=t", "-e=x"]) == 0
    assert cli.main(["snarkjs", "zkey", "export", "verificationkey", f"{d}/c_final.zkey", f"{d}/vkey.json"]) == 0
    json.dump(I.secure_agg_client_input(), open(f"{d}/input.json", "w"))
    wasm = f"{d}/secure_agg_client_js/secure_agg_client.wasm"
    assert cli.main(["snarkjs", "wtns", "calculate", wasm, f"{d}/input.json", f"{d}/w.wtns"]) == 0
    assert cli.main(["generate_witness", wasm, f"{d}/input.json", f"{d}/w2.wtns"]) == 0
    assert open(f"{d}/w.wtns", "rb").read() == open(f"{d}/w2.wtns", "rb").read()
    assert cli.main(["snarkjs", "groth16", "prove", f"{d}/c_final.zkey", f"{d}/w.wtns", f"{d}/proof.json", f"{d}/public.json"]) == 0
    assert cli.main(["snarkjs", "groth16", "verify", f"{d}/vkey.json", f"{d}/public.json", f"{d}/proof.json"]) == 0
    pub = json.load(open(f"{d}/public.json"))
    bad = [str(int(pub[0]) + 1)] + pub[1:]
    json.dump(bad, open(f"{d}/public_bad.json", "w"))
    assert cli.main(["snarkjs", "groth16", "verify", f"{d}/vkey.json", f"{d}/public_bad.json", f"{d}/proof.json"]) == 1
    bad_in = dict(I.secure_agg_client_input(), tau_squared="0", gradient=["3"] + ["0"] * 7)
    json.dump(bad_in, open(f"{d}/bad_input.json", "w"))
    assert cli.main(["snarkjs", "wtns", "calculate", wasm, f"{d}/bad_input.json", f"{d}/bad.wtns"]) == 1   # Assert Failed
    # fixed blinding: API result equals the oracle's proof bytes; the oracle's pairing accepts it too
    zk = open(f"{d}/c_final.zkey", "rb").read()
    res = sj.groth16.fullProve(I.secure_agg_client_input(), wasm, f"{d}/c_final.zkey", rs=(5, 6))
    w = formats.wtns_read(open(f"{d}/w.wtns", "rb").read())
    ref_proof, ref_pub = ol.groth16_prove(zk, w, 5, 6)
    assert formats.proof_json_to_bytes(res["proof"]) == ref_proof
    assert res["publicSignals"] == [str(x) for x in ol.ints(ref_pub)]
    vk = json.load(open(f"{d}/vkey.json"))
    assert sj.groth16.verify(vk, res["publicSignals"], res["proof"])
    assert g16.verify(g16.vkey_from_json(vk), ol.ints(ref_pub), g16.proof_from_json(res["proof"]))


def test_product_setup_matches_the_oracle_setup(emul_prover, oracle):
    """`groth16 setup` replacement (host scalar arithmetic + device generator multiplications) against the oracle's own
    setup from the same toxic waste: every point section and the header byte-identical, coefficient sections equal as sets
    (snarkjs interleaves A/B per constraint, we write A then B: the prover does not depend on the order)."""
    import struct
    import groth16_ref as g16
    import witness_ref as wr
    from zkfl_b200.zkey_setup import toxic_from_seed
    cc = build_circuit("secure_agg_client")
    mine = emul_prover.new_zkey(cc, b"parity-seed")
    ref = g16.setup_fast(wr.R1cs(cc.r1cs_bytes()), *toxic_from_seed(b"parity-seed"))
    _, a = wr.read_sections(mine, b"zkey")
    _, b = wr.read_sections(ref, b"zkey")
    for sid in (1, 2, 3, 5, 6, 7, 8, 9):
        assert a[sid] == b[sid], f"section {sid} differs"

    def coeff_set(sec):
        n = struct.unpack_from("<I", sec, 0)[0]
        return {sec[4 + 44 * i:48 + 44 * i] for i in range(n)}
    assert coeff_set(a[4]) == coeff_set(b[4]) and len(a[4]) == len(b[4])


def test_c_abi_argument_and_format_errors(emul_prover):
    """errors come back as codes + messages, never as crashes: mismatched artefacts, truncated files, unreduced values"""
    import ctypes
    from zkfl_b200.api import Circuit
    lib, ctx = emul_prover.lib, emul_prover.ctx
    cc = pc.tiny_circuit()
    circ = emul_prover.load_circuit(cc)
    other = emul_prover.load_circuit(build_circuit("secure_agg_client"), check_constraints=False)
    zk = open(os.path.join(os.path.dirname(__file__), "golden", "tiny.zkey"), "rb").read()
    Z = emul_prover.load_zkey(zk)
    with pytest.raises(_lib.ZkflError, match="do not match"):
        emul_prover.full_prove(other, Z, [I.secure_agg_client_input()], [(1, 2)], check=False)
    with pytest.raises(ValueError, match="r1cs"):      # a constraint check that cannot be done is refused, never skipped
        emul_prover.full_prove(other, Z, [I.secure_agg_client_input()], [(1, 2)])
    with pytest.raises(_lib.ZkflError, match="not reduced"):
        emul_prover.full_prove(circ, Z, pc.tiny_inputs()[:1], [(2 ** 255, 1)])
    with pytest.raises(_lib.ZkflError):
        emul_prover.load_zkey(zk[:len(zk) // 2])
    with pytest.raises(_lib.ZkflError):
        emul_prover.load_zkey(b"wtns" + zk[4:])
    bad = bytearray(cc.program_bytes())
    bad[12 + 12 + 12] ^= 0xFF   # n_ops in the header no longer matches the op section
    with pytest.raises(_lib.ZkflError):
        Circuit(emul_prover, bytes(bad))
    # every index of a program loaded from disk is validated at load time (ops, LC ids, LC offsets, term wires, Poseidon inputs)
    import struct
    from zkfl_b200.formats import read_container, write_container
    secs = read_container(cc.program_bytes(), b"zkwp")
    hdr = struct.unpack("<8I", secs[1])
    n_wires, n_ops, n_lcs, n_terms, n_pos = hdr[0], hdr[3], hdr[4], hdr[5], hdr[6]

    def mutated(sid, off, value):
        m = dict(secs)
        b = bytearray(m[sid]); b[off:off + 4] = struct.pack("<I", value); m[sid] = bytes(b)
        return write_container(b"zkwp", 1, sorted(m.items()))
    ops = struct.unpack(f"<{5 * n_ops}I", secs[2])
    first = {code: next(o for o in range(n_ops) if ops[5 * o] == code) for code in (1, 2, 3, 4) if code in ops[0::5]}
    cases = [(3, 4, n_terms + 1),                                  # LC offsets not monotone
             (4, 0, n_wires),                                      # term wire out of range
             (6, 0, n_wires)]                                      # Poseidon input wire out of range
    for code, o in first.items():
        cases.append((2, 20 * o + 4, n_wires))                     # dst out of range
        cases.append((2, 20 * o + 4, 0))                           # dst inside the inputs
        if code <= 3:
            cases.append((2, 20 * o + 8, n_lcs))                   # LC id out of range
        if code == 2:
            cases.append((2, 20 * o + 12, n_lcs + 7))
        if code == 3:
            cases.append((2, 20 * o + 12, 255))                    # more bits than the field has
            cases.append((2, 20 * o + 4, n_wires - 1))             # dst + n_bits past the end
        if code == 4:
            cases.append((2, 20 * o + 12, n_pos))                  # Poseidon input list past the end
            cases.append((2, 20 * o + 8, 18))                      # width without constants
            cases.append((2, 20 * o + 4, n_wires - 5))             # outputs past the end
    assert set(first) >= {2, 3, 4}
    for sid, off, value in cases:
        with pytest.raises(_lib.ZkflError):
            Circuit(emul_prover, mutated(sid, off, value))
    Circuit(emul_prover, write_container(b"zkwp", 1, sorted(secs.items()))).close()   # the unmutated container still loads
    # a witness handed to `groth16 prove` must be well-formed: every element reduced, wire 0 == 1
    good = emul_prover.calculate_witness(circ, pc.tiny_inputs()[:1])[0]
    emul_prover.prove(Z, [good], [(1, 2)])
    with pytest.raises(_lib.ZkflError, match="wire 0"):
        emul_prover.prove(Z, [(2).to_bytes(32, "little") + good[32:]], [(1, 2)])
    with pytest.raises(_lib.ZkflError, match="not reduced"):
        emul_prover.prove(Z, [good[:64] + b"\xff" * 32 + good[96:]], [(1, 2)])
    # fullProve aborts on a failed === (one pass, check on the device) and returns no proof; unreduced inputs are refused
    with pytest.raises(_lib.AssertFailed):
        emul_prover.full_prove(circ, Z, [pc.tiny_inputs()[0], {"out": "5", "bound": "17", "x": "3", "y": "5"}], [(1, 2), (3, 4)])
    bad_first = (ctypes.c_uint32 * 2)()
    packed = circ.pack_inputs([pc.tiny_inputs()[0], {"out": "5", "bound": "17", "x": "3", "y": "5"}])
    out_p = ctypes.create_string_buffer(b"\x01" * 512, 512)
    rc = lib.zkfl_groth16_full_prove_batch(ctx, circ.handle, Z.handle, circ.r1cs_handle, _lib.as_ptr(packed), None, 2, out_p, None, bad_first)
    assert rc == -5 and bad_first[0] == 0xFFFFFFFF and bad_first[1] != 0xFFFFFFFF and out_p.raw == bytes(512)
    with pytest.raises(_lib.ZkflError, match="not reduced"):
        emul_prover.full_prove(circ, Z, b"\xff" * (32 * circ.n_inputs), [(1, 2)])
    # an asynchronous run whose verdicts nobody fetched must not leak into the next, smaller call (it once overflowed first_bad)
    three = circ.pack_inputs(pc.tiny_inputs())
    emul_prover.stage(circ, Z, three, emul_prover._pack_rs([(1, 2), (3, 4), (5, 6)], 3), 3)
    emul_prover.run_staged(circ, Z, 3, check=True)
    guard = (ctypes.c_uint32 * 4)(7, 7, 7, 7)
    one = circ.pack_inputs(pc.tiny_inputs()[:1])
    p1 = ctypes.create_string_buffer(256)
    assert lib.zkfl_groth16_full_prove_batch(ctx, circ.handle, Z.handle, circ.r1cs_handle, _lib.as_ptr(one), None, 1, p1, None, guard) == 0
    assert list(guard) == [0xFFFFFFFF, 7, 7, 7]
    h = ctypes.c_void_p()
    assert lib.zkfl_zkey_load(ctx, None, 0, ctypes.byref(h)) != 0 and lib.zkfl_last_error()
    assert lib.zkfl_groth16_prove_batch(ctx, Z.handle, None, None, 1, None, None) != 0
    circ.close(); other.close(); Z.close()


def test_commitment_pipeline_matches_the_reference_helpers(emul_prover):
    """the off-circuit commitments (Merkle tree, root_D/root_W/root_G/root_K, PRF masks) computed by the batched evaluator
    equal the oracle's restatement of the reference's JavaScript helpers for the simulation's clients"""
    import bn254_ref as bn
    from zkfl_b200 import commitments
    clients = I.simulation_clients(3)
    req = []
    for c in clients:
        c.training_input([3, -2, 0, 7])
        peers = [j for j in (1, 2, 3) if j != c.id]
        req.append({"features": c.features, "labels": c.labels, "weights": c.weights, "gradient": c.gradient, "client_id": c.id,
                    "round": c.ROUND, "master_key": bn.poseidon([c.id, 12345]), "peer_ids": peers,
                    "shared_keys": [bn.poseidon([min(c.id, j), max(c.id, j), 12345]) for j in peers]})
    got = commitments.compute(emul_prover, req, 8, 4, 3)
    for c, r, g in zip(clients, req, got):
        tree = bn.build_merkle_tree([bn.vector_hash(f + [l]) for f, l in zip(c.features, c.labels)], 3)
        assert g["tree"] == tree and g["root_D"] == c.root_d
        assert g["root_W"] == bn.weight_commitment(c.weights) == c.root_w
        assert g["root_G"] == bn.gradient_commitment([x % bn.R for x in c.gradient], c.id, c.ROUND) == c.root_g
        assert g["root_K"] == bn.key_material_commitment(r["master_key"], r["shared_keys"])
        for j, key, mask in zip(r["peer_ids"], r["shared_keys"], g["masks"]):
            assert mask == bn.derive_pairwise_mask(key, c.ROUND, c.id, j, 4)


def test_setup_entropy_and_contribution(emul_prover, oracle):
    """keys must not be derivable.  Default setup draws OS randomness (two setups differ); `zkey contribute`
    re-randomises delta (delta1/delta2, C and H sections change, A/B sections and IC stay) and appends a contribution record;
    proofs under the contributed key verify under ITS verification key and not under the old one."""
    import groth16_ref as g16
    import witness_ref as wr
    from zkfl_b200 import formats
    cc = pc.tiny_circuit()
    k1, k2 = emul_prover.new_zkey(cc), emul_prover.new_zkey(cc)
    assert k1 != k2
    base = emul_prover.new_zkey(cc, b"contrib-test")
    assert base == emul_prover.new_zkey(cc, b"contrib-test")          # explicit seed: reproducible (tests only)
    c1 = emul_prover.contribute_zkey(base, "alice", b"e1")
    c2 = emul_prover.contribute_zkey(base, "alice", b"e1")
    assert c1 != base and c1 != c2                                     # fresh CSPRNG secret each time
    _, a = wr.read_sections(base, b"zkey")
    _, b = wr.read_sections(c1, b"zkey")
    for sid in (1, 3, 4, 5, 6, 7):
        assert a[sid] == b[sid]
    assert a[8] != b[8] and a[9] != b[9] and a[2] != b[2] and a[2][:84 + 64 + 64 + 128 + 128] == b[2][:84 + 64 + 64 + 128 + 128]
    import struct
    assert struct.unpack_from("<I", b[10], 64)[0] == 1 and b"alice" in b[10]
    c12 = emul_prover.contribute_zkey(c1, "bob")
    assert struct.unpack_from("<I", wr.read_sections(c12, b"zkey")[1][10], 64)[0] == 2
    circ = emul_prover.load_circuit(cc)
    Z = emul_prover.load_zkey(c12)
    proofs, pubs = emul_prover.full_prove(circ, Z, pc.tiny_inputs()[:1], [(5, 6)])
    w = emul_prover.calculate_witness(circ, pc.tiny_inputs()[:1])[0]
    assert (proofs[0], pubs[0]) == oracle.groth16_prove(c12, w, 5, 6)
    vk_new = g16.vkey_from_json(formats.export_verification_key(c12))
    vk_old = g16.vkey_from_json(formats.export_verification_key(base))
    assert g16.verify(vk_new, oracle.ints(pubs[0]), g16.proof_from_bytes(proofs[0]))
    assert not g16.verify(vk_old, oracle.ints(pubs[0]), g16.proof_from_bytes(proofs[0]))
    Z.close(); circ.close()


def test_proof_json_decoding_is_strict():
    """proof.json must say groth16 / bn128 and carry affine points (z = 1)"""
    from zkfl_b200 import formats
    pj = formats.proof_bytes_to_json(bytes(range(1, 33)) * 8)
    assert formats.proof_json_to_bytes(pj) == bytes(range(1, 33)) * 8
    for bad in (dict(pj, protocol="plonk"), dict(pj, curve="bls12381"), dict(pj, pi_a=pj["pi_a"][:2] + ["2"]),
                dict(pj, pi_b=pj["pi_b"][:2] + [["0", "1"]]), dict(pj, pi_c=pj["pi_c"][:2])):
        with pytest.raises((ValueError, IndexError)):
            formats.proof_json_to_bytes(bad)


def _reference_aggregate(masked, accept, model, lr):
    """restatement of Server.aggregateUpdates (tests/full_system_simulation.mjs:1137-1199) with JavaScript's semantics: BigInt sums
    mod p, `> p / 2n` means negative, Number(BigInt) (round to nearest even), then IEEE-double mean and model update"""
    import bn254_ref as bn
    ids = [i for i, a in enumerate(accept) if a]
    if not ids:
        return None
    dim = len(model)
    agg = [sum(masked[i][j] for i in ids) % bn.R for j in range(dim)]
    g = [float(a - bn.R) if a > bn.R // 2 else float(a) for a in agg]
    g = [x / len(ids) for x in g]
    return {"aggregated_field": agg, "aggregated_gradient": g, "new_model": [m - lr * x for m, x in zip(model, g)], "num_clients": len(ids)}


def case_aggregate(P):
    import random
    import bn254_ref as bn
    rnd = random.Random(3)
    for n, dim in ((3, 4), (200, 4), (1023, 7)):
        grads = [[rnd.randrange(-10 ** 6, 10 ** 6) for _ in range(dim)] for _ in range(n)]
        masks = [[rnd.randrange(bn.R) for _ in range(dim)] for _ in range(n)]
        masked = [[(grads[i][j] + masks[i][j] - masks[(i + 1) % n][j]) % bn.R for j in range(dim)] for i in range(n)]
        model = [rnd.uniform(-1, 1) for _ in range(dim)]
        got = P.aggregate_updates(masked, [True] * n, model, 0.01)
        assert got == _reference_aggregate(masked, [True] * n, model, 0.01)
        assert got["aggregated_gradient"] == [sum(g[j] for g in grads) / n for j in range(dim)]      # the masks cancel
        accept = [rnd.random() < 0.7 for _ in range(n)]      # masks no longer cancel: 254-bit magnitudes through the double conversion
        assert P.aggregate_updates(masked, accept, model, 0.01) == _reference_aggregate(masked, accept, model, 0.01)
    assert P.aggregate_updates([[1, 2]], [False], [0.0, 0.0], 0.01) is None
    with pytest.raises(_lib.ZkflError, match="not reduced"):
        P.aggregate_updates([[bn.R, 0]], [True], [0.0, 0.0], 0.01)


def test_masked_aggregation_and_model_update(emul_prover):
    """SURVEY 8f item 4 / VERDICT r1 item 7: aggregateUpdates as device kernels against the restated JavaScript"""
    case_aggregate(emul_prover)


def test_gpu_commitment_inputs_equal_the_host_generators(emul_prover):
    """clients whose commitments come from the batched commitment program (commitments.hydrate: trees, roots, derived key
    material, masks) produce exactly the inputs of the per-client host generators, for all three circuits"""
    from zkfl_b200 import commitments
    ref = I.simulation_clients(6)
    lcg = I.JsLcg(12345)
    mine = [I.SimClient(i, lcg, hashed=False) for i in range(1, 7)]
    commitments.hydrate(emul_prover, mine, [3, -2, 0, 7])
    for a, b in zip(ref, mine):
        assert a.balance_input() == b.balance_input()
        assert a.training_input([3, -2, 0, 7]) == b.training_input([3, -2, 0, 7])
        base = 3 * ((a.id - 1) // 3)
        peers = [base + k for k in (1, 2, 3) if base + k != a.id]
        assert a.secagg_input(peers) == b.secagg_input(peers)
