"""One federated round on the GPU backend: the flow of the reference's system test
(tests/full_system_simulation.mjs:1244-1395 `runSimulation`), batch-first.

The reference proves client by client through child processes (three per proof). Here every phase proves ALL clients
of the round in one batched GPU call per circuit (independent proofs, shared bases), then the server side performs the
reference's checks: public-signal positions and values (:848-880, :886-990, :995-1131), cross-proof bindings
(root_D / root_G / root_W), recomputation of root_G from the clear gradient, `groth16 verify`, and the masked
aggregation with the model update (:1137-1199). Clients are grouped in federations of three (NUM_PEERS = 2, as the
`SecureMaskedUpdate(4, 2)` main component requires).
"""
from __future__ import annotations

import time

from . import formats, inputs
from . import snarkjs as sj
from .circuits import build_circuit
from .circuits.poseidon_params import FR

LEARNING_RATE = 0.01  # full_system_simulation.mjs:52


def _setup(prover, name, cache, setup_seed: bytes | None = None):
    """per-circuit setup; the toxic waste is OS randomness unless a test passes `setup_seed` (reproducible, forgeable)"""
    if name not in cache:
        cc = build_circuit(name)
        zk = prover.new_zkey(cc, None if setup_seed is None else setup_seed + name.encode())
        cache[name] = (prover.load_circuit(cc), prover.load_zkey(zk), formats.export_verification_key(zk))
    return cache[name]


def run_round(prover, n_clients: int = 3, seed: int = 12345, weights=None, verify: bool = True, cache: dict | None = None,
              setup_seed: bytes | None = None) -> dict:
    assert n_clients % 3 == 0, "clients come in federations of three (NUM_PEERS = 2)"
    cache = {} if cache is None else cache
    t0 = time.perf_counter()
    clients = inputs.simulation_clients(n_clients, seed)                       # phases 1-2: data, Merkle root_D
    model = list(weights) if weights is not None else [0] * clients[0].DIM     # Server.initializeModel (:817-822)
    timing = {"inputs_s": time.perf_counter() - t0}
    report = {"clients": n_clients, "verified": {"balance": 0, "training": 0, "secagg": 0}}

    def prove_phase(name, ins):
        circ, zkey, vk = _setup(prover, name, cache, setup_seed)
        t = time.perf_counter()
        proofs, pubs = prover.full_prove(circ, zkey, ins)                      # constraint check inside the pass; r, s from the OS like snarkjs
        timing[name + "_s"] = time.perf_counter() - t
        return vk, [formats.proof_bytes_to_json(p) for p in proofs], [formats.publics_bytes_to_json(q) for q in pubs]

    def verify_all(vk, sigs, proofs):
        if not verify:
            return [True] * len(sigs)
        t = time.perf_counter()
        ok = sj.groth16.verifyBatch(vk, list(zip(sigs, proofs)), prover=prover)   # one GPU pass per phase
        timing["verify_s"] = timing.get("verify_s", 0.0) + time.perf_counter() - t
        return ok

    # phase 3: balance proofs; Server.verifyBalanceProof (:848-880)
    vk, proofs, sigs = prove_phase("balance_unified", [c.balance_input() for c in clients])
    balance_root = {}
    valid = verify_all(vk, sigs, proofs)
    for c, p, s, v in zip(clients, proofs, sigs, valid):
        ok = s[1] == str(c.root_d) and s[2] == str(c.N) and int(s[3]) + int(s[4]) == c.N and v
        if ok:
            balance_root[c.id] = s[1]
            report["verified"]["balance"] += 1

    # phase 4: verified-gradient proofs; Server.verifyTrainingProof (:886-990)
    vk, proofs, sigs = prove_phase("sgd_verified", [c.training_input(model) for c in clients])
    trained = set()
    valid = verify_all(vk, sigs, proofs)
    for c, p, s, v in zip(clients, proofs, sigs, valid):
        ok = (balance_root.get(c.id) == str(c.root_d)                                           # binding to the balance proof
              and s[1] == str(c.ROUND) and s[2] == str(c.root_d) and s[3] == str(c.root_g) and s[4] == str(c.root_w)
              and s[5] == str(c.TAU2)
              and inputs.gradient_commitment(c.gradient, c.id, c.ROUND) == c.root_g            # recomputed from the clear gradient
              and v)
        if ok:
            trained.add(c.id)
            report["verified"]["training"] += 1

    # phase 4.5: secure aggregation proofs; Server.verifySecureAggregationProof (:995-1131)
    def peers_of(c):
        base = 3 * ((c.id - 1) // 3)
        return [base + k for k in (1, 2, 3) if base + k != c.id]

    vk, proofs, sigs = prove_phase("secure_masked_update", [c.secagg_input(peers_of(c)) for c in clients])
    accepted = []
    valid = verify_all(vk, sigs, proofs)
    for c, p, s, v in zip(clients, proofs, sigs, valid):
        ok = (c.id in trained and s[0] == str(c.id) and s[1] == str(c.ROUND) and s[2] == str(c.root_d) and s[3] == str(c.root_g)
              and s[4] == str(c.root_w) and s[6] == str(c.TAU2) and s[7:11] == [str(x) for x in c.masked_update]
              and s[11:13] == [str(j) for j in peers_of(c)] and v)
        if ok:
            accepted.append(c)
            report["verified"]["secagg"] += 1

    # phase 5: Server.aggregateUpdates (:1137-1199)
    agg = [sum(c.masked_update[k] for c in accepted) % FR for k in range(clients[0].DIM)]
    signed = [a - FR if a > FR // 2 else a for a in agg]
    mean = [g / max(len(accepted), 1) for g in signed]
    report["aggregated_gradient"] = mean
    report["expected_gradient"] = [sum(c.gradient[k] for c in accepted) / max(len(accepted), 1) for k in range(clients[0].DIM)]
    report["new_model"] = [w - LEARNING_RATE * g for w, g in zip(model, mean)]
    report["timing"] = timing
    report["proofs"] = 3 * n_clients
    return report


if __name__ == "__main__":
    import argparse
    import json
    from .api import Prover
    ap = argparse.ArgumentParser()
    ap.add_argument("--clients", type=int, default=3)
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--repeat", type=int, default=1, help="run the round this many times on one prover (keys cached) and print the last")
    ap.add_argument("--brief", action="store_true", help="print counts and timings only")
    a = ap.parse_args()
    P, cache, rep = Prover(0), {}, None
    for _ in range(max(1, a.repeat)):
        rep = run_round(P, a.clients, verify=not a.no_verify, cache=cache)
    if a.brief:
        rep = {k: rep[k] for k in ("clients", "proofs", "verified", "timing")}
    print(json.dumps(rep, indent=1))
