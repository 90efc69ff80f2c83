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


def _dist():
    """(torch.distributed module, rank, world) when a process group with more than one rank is up, else (None, 0, 1)"""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist, dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return None, 0, 1


def _setup(prover, name, cache, setup_seed: bytes | None = None):
    """per-circuit setup; the toxic waste is OS randomness unless a test passes `setup_seed` (reproducible, forgeable).
    With several ranks, rank 0 makes the key and broadcasts the bytes: every GPU must prove under the SAME key."""
    if name not in cache:
        dist, rank, world = _dist()
        cc = build_circuit(name)
        zk = prover.new_zkey(cc, None if setup_seed is None else setup_seed + name.encode()) if rank == 0 else None
        if dist is not None:
            box = [zk]
            dist.broadcast_object_list(box, src=0)
            zk = box[0]
        cache[name] = (prover.load_circuit(cc), prover.load_zkey(zk), formats.export_verification_key(zk))
    return cache[name]


def run_round(prover, n_clients: int = 3, seed: int = 12345, weights=None, verify: bool = True, cache: dict | None = None,
              setup_seed: bytes | None = None, gpu_inputs: bool = True) -> dict | None:
    """One round.  With a torch.distributed process group of G > 1 ranks (one per GPU) the proofs of every phase are sharded
    b -> rank b mod G (sharding.prove_independent, no data-path collective) and rank 0 plays the server (verification, aggregation);
    the other ranks return None.  gpu_inputs: the clients' commitments (Merkle trees, roots, key material, masks) come from the GPU
    commitment pipeline in one batched pass instead of per-client Poseidon evaluations on the host."""
    assert n_clients % 3 == 0, "clients come in federations of three (NUM_PEERS = 2)"
    from . import commitments, sharding
    dist, rank, world = _dist()
    cache = {} if cache is None else cache
    t0 = time.perf_counter()
    lcg = inputs.JsLcg(seed)                                                   # phases 1-2: data, Merkle root_D
    clients = [inputs.SimClient(i, lcg, hashed=not gpu_inputs) for i in range(1, n_clients + 1)]
    model = list(weights) if weights is not None else [0] * clients[0].DIM     # Server.initializeModel (:817-822)
    if gpu_inputs:
        commitments.hydrate(prover, clients, model)
    timing = {"inputs_s": time.perf_counter() - t0}
    report = {"clients": n_clients, "verified": {"balance": 0, "training": 0, "secagg": 0}, "n_gpus": world}

    def prove_phase(name, ins):
        circ, zkey, vk = _setup(prover, name, cache, setup_seed)
        t = time.perf_counter()
        if dist is not None:
            proofs, pubs = sharding.prove_independent(prover, circ, zkey, ins)  # constraint check inside every rank's pass
        else:
            proofs, pubs = prover.full_prove(circ, zkey, ins)                  # constraint check inside the pass; r, s from the OS like snarkjs
        timing[name + "_s"] = time.perf_counter() - t
        if rank != 0:
            return vk, None, None
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
    if rank == 0:
        valid = verify_all(vk, sigs, proofs)
        for c, p, s, v in zip(clients, proofs, sigs, valid):
            ok = s[1] == str(c.root_d) and s[2] == str(c.N) and int(s[3]) + int(s[4]) == c.N and v
            if ok:
                balance_root[c.id] = s[1]
                report["verified"]["balance"] += 1

    # phase 4: verified-gradient proofs; Server.verifyTrainingProof (:886-990)
    vk, proofs, sigs = prove_phase("sgd_verified", [c.training_input(model) for c in clients])
    trained = set()
    if rank == 0:
        valid = verify_all(vk, sigs, proofs)
        for c, p, s, v in zip(clients, proofs, sigs, valid):
            # root_G recomputed from the clear gradient: c.root_g (GPU commitment pipeline or host Poseidon, both from c.gradient)
            recomputed = c.root_g if gpu_inputs else inputs.gradient_commitment(c.gradient, c.id, c.ROUND)
            ok = (balance_root.get(c.id) == str(c.root_d)                                       # binding to the balance proof
                  and s[1] == str(c.ROUND) and s[2] == str(c.root_d) and s[3] == str(recomputed) and s[4] == str(c.root_w)
                  and s[5] == str(c.TAU2) and v)
            if ok:
                trained.add(c.id)
                report["verified"]["training"] += 1

    # phase 4.5: secure aggregation proofs; Server.verifySecureAggregationProof (:995-1131)
    def peers_of(c):
        base = 3 * ((c.id - 1) // 3)
        return [base + k for k in (1, 2, 3) if base + k != c.id]

    vk, proofs, sigs = prove_phase("secure_masked_update", [c.secagg_input(peers_of(c)) for c in clients])
    if rank != 0:
        return None
    accepted = []
    valid = verify_all(vk, sigs, proofs)
    for c, p, s, v in zip(clients, proofs, sigs, valid):
        ok = (c.id in trained and s[0] == str(c.id) and s[1] == str(c.ROUND) and s[2] == str(c.root_d) and s[3] == str(c.root_g)
              and s[4] == str(c.root_w) and s[6] == str(c.TAU2) and s[7:11] == [str(x) for x in c.masked_update]
              and s[11:13] == [str(j) for j in peers_of(c)] and v)
        accepted.append(bool(ok))
        if ok:
            report["verified"]["secagg"] += 1

    # phase 5: Server.aggregateUpdates (:1137-1199) on the device: field sum, signed decode, mean, SGD step
    t = time.perf_counter()
    agg = prover.aggregate_updates([c.masked_update for c in clients], accepted, model, LEARNING_RATE)
    timing["aggregate_s"] = time.perf_counter() - t
    n_acc = sum(accepted)
    report["aggregated_gradient"] = agg["aggregated_gradient"] if agg else None
    report["expected_gradient"] = [sum(c.gradient[k] for c, a in zip(clients, accepted) if a) / max(n_acc, 1) for k in range(clients[0].DIM)]
    report["new_model"] = agg["new_model"] if agg else None
    report["timing"] = timing
    report["timing"]["round_s"] = time.perf_counter() - t0
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
    ap.add_argument("--host-inputs", action="store_true", help="client commitments by per-client Poseidon on the host (the slow path)")
    a = ap.parse_args()
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:                         # torchrun --nproc-per-node G -m zkfl_b200.simulation ...: one process per GPU, NCCL
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, cache, rep = Prover(local), {}, None
    for _ in range(max(1, a.repeat)):
        rep = run_round(P, a.clients, verify=not a.no_verify, cache=cache, gpu_inputs=not a.host_inputs)
    if rep is not None:
        if a.brief:
            rep = {k: rep[k] for k in ("clients", "n_gpus", "proofs", "verified", "timing")}
        print(json.dumps(rep, indent=1))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
