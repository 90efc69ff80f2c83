"""Circuit front-end: the restated circuits accept exactly what the reference's generators produce.
Witnesses come from the C++ oracle interpreter, constraints are checked by the Python oracle on the
`.r1cs` bytes (the reference's tests only check exit codes and public signals: SURVEY section 4)."""
import json
import os

import pytest

import bn254_ref as bn
import witness_ref as wr
import zkfl_b200  # noqa: F401
from zkfl_b200 import inputs as I
from zkfl_b200.circuits import CIRCUIT_NAMES, build_circuit

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
_R1CS = {}


def witness_and_check(oracle, name, inp):
    cc = build_circuit(name)
    w = oracle.ints(oracle.witness_batch(cc.program_bytes(), oracle.fes(cc.flatten_input(inp)), cc.n_inputs, cc.n_wires))
    if name not in _R1CS:
        _R1CS[name] = wr.R1cs(cc.r1cs_bytes())
    return w, _R1CS[name].first_violation(w)


def test_circuit_census():
    """sizes of the restated circuits (multiplicative constraints + the linear ones we keep); the
    domain sizes match SURVEY section 8's census except secure_masked_update (2^13, flagged 'tight' there)."""
    expect_domain = {"balance_unified": 14, "sgd_verified": 14, "sgd_step_quick": 14, "sgd_step_v5": 15,
                     "secure_masked_update": 13, "secure_agg_client": 12}
    expect_public = {"balance_unified": 5, "sgd_verified": 6, "sgd_step_quick": 5, "sgd_step_v5": 5,
                     "secure_masked_update": 13, "secure_agg_client": 12}
    for name, lg in expect_domain.items():
        cc = build_circuit(name)
        assert cc.n_public == expect_public[name]
        assert (cc.n_constraints + cc.n_public).bit_length() == lg, (name, cc.n_constraints)


def test_v5_fixture_satisfies_sgd_step_v5(oracle):
    d = json.load(open(os.path.join(GOLDEN, "test_input_v5.json")))
    w, bad = witness_and_check(oracle, "sgd_step_v5", d)
    assert bad is None
    # public signal order [client_id, round, root_D, root_G, tauSquared] (sgd_step_v5.circom:168)
    assert w[1:6] == [1, 1, int(d["root_D"]), int(d["root_G"]), int(d["tauSquared"])]
    # python interpreter agrees with the C++ one
    cc = build_circuit("sgd_step_v5")
    assert wr.calculate_witness(wr.Program(cc.program_bytes()), cc.flatten_input(d)) == w
    for key, delta in (("root_G", 1), ("root_D", 1), ("tauSquared", -70000)):
        t = dict(d)
        t[key] = str(int(d[key]) + delta)
        assert witness_and_check(oracle, "sgd_step_v5", t)[1] is not None, key


def test_full_system_simulation_clients(oracle):
    clients = I.simulation_clients(3)
    for c in clients:
        w, bad = witness_and_check(oracle, "balance_unified", c.balance_input())
        assert bad is None and w[1:6] == [c.id, c.root_d, 8, c.c0, c.c1]
        w, bad = witness_and_check(oracle, "sgd_verified", c.training_input([0] * 4))
        assert bad is None and w[1:7] == [c.id, 1, c.root_d, c.root_g, c.root_w, c.TAU2]
        peers = [j for j in (1, 2, 3) if j != c.id]
        w, bad = witness_and_check(oracle, "secure_masked_update", c.secagg_input(peers))
        assert bad is None and w[8:12] == c.masked_update and w[12:14] == peers
        # commitments recomputed by the oracle's own helpers
        assert c.root_d == bn.build_merkle_tree([bn.vector_hash(f + [l]) for f, l in zip(c.features, c.labels)], 3)[-1][0]
        assert c.root_g == bn.gradient_commitment([g % bn.R for g in c.gradient], c.id, 1)
        assert c.root_w == bn.weight_commitment(c.weights)
    for k in range(4):  # masks cancel (test_secure_aggregation.mjs:215-238)
        assert sum(c.masked_update[k] for c in clients) % bn.R == sum(c.gradient[k] for c in clients) % bn.R


def test_negative_field_inputs_and_other_circuits(oracle):
    for inp in I.sgd_verified_batch(2, nonzero_weights=True):
        assert any(int(x) < 0 for x in inp["weights"] + inp["expectedSummedGrad"])
        assert witness_and_check(oracle, "sgd_verified", inp)[1] is None
    assert witness_and_check(oracle, "secure_agg_client", I.secure_agg_client_input())[1] is None
    assert witness_and_check(oracle, "balance_unified", I.balance_integration_input())[1] is None
    q = I.simulation_clients(1)[0].training_input([0] * 4)
    for k in ("weights", "expectedSummedGrad", "remainder", "root_W"):
        q.pop(k)
    assert witness_and_check(oracle, "sgd_step_quick", q)[1] is None


def test_bad_inputs_are_rejected(oracle):
    c = I.simulation_clients(1)[0]
    inp = c.training_input([0] * 4)
    bad = dict(inp)
    bad["gradPos"] = ["1"] + inp["gradPos"][1:]          # wrong gradient
    assert witness_and_check(oracle, "sgd_verified", bad)[1] is not None
    bad = dict(inp)
    bad["remainder"] = [str(int(inp["remainder"][0]) + 8000)] + inp["remainder"][1:]   # remainder >= divisor
    assert witness_and_check(oracle, "sgd_verified", bad)[1] is not None
    cc = build_circuit("sgd_verified")
    with pytest.raises(KeyError):
        cc.flatten_input({k: v for k, v in inp.items() if k != "round"})
    with pytest.raises(ValueError):
        cc.flatten_input({**inp, "weights": inp["weights"][:3]})


def test_r1cs_and_program_round_trip():
    cc = build_circuit("secure_agg_client")
    r1 = wr.R1cs(cc.r1cs_bytes())
    assert (r1.n_wires, r1.n_constraints, r1.n_public) == (cc.n_wires, cc.n_constraints, cc.n_public)
    prog = wr.Program(cc.program_bytes())   # also asserts the embedded Poseidon constants equal the oracle's
    assert prog.n_wires == cc.n_wires and prog.meta["inputs"][0]["name"] == "client_id"
    assert set(CIRCUIT_NAMES) >= {"balance_unified", "balance_unified_prod", "sgd_verified", "secure_masked_update"}
