"""Off-circuit commitment pipeline on the GPU.

The reference computes every commitment a client publishes -- leaf hashes and Merkle tree of the dataset (root_D and the
authentication paths), weight / gradient / key-material commitments, pairwise PRF masks -- on the CPU with circomlibjs
(tests/full_system_simulation.mjs:139-238, 308-335, 434-438, 567-609).  Here they are the wires of a small *commitment
program* written with the same front-end as the circuits and run for ALL clients at once by the batched witness
evaluator (`zkfl_wtns_eval_wires`); no new kernels are involved.
"""
from __future__ import annotations

from functools import lru_cache

from .circuits import templates as T
from .circuits.builder import CircuitBuilder
from .circuits.poseidon_params import FR


@lru_cache(maxsize=None)
def commitment_program(n: int, dim: int, depth: int, num_peers: int = 2, derive_keys: bool = False):
    """Returns (compiled program, wire map). Inputs per client: features[n][dim], labels[n], weights[dim], gradient[dim],
    client_id, round, peer_lo[num_peers], peer_hi[num_peers] (min/max of the id pair) and either master_key +
    shared_keys[num_peers], or (derive_keys) key_seed from which the program derives masterKey = Poseidon(id, seed) and
    K_ij = Poseidon(min, max, seed) itself (tests/full_system_simulation.mjs:1321-1336)."""
    assert n == 1 << depth
    c = CircuitBuilder(f"commitments_{n}_{dim}_{depth}_{num_peers}_{int(derive_keys)}")
    features = c.input("features", (n, dim))
    labels = c.input("labels", (n,))
    weights = c.input("weights", (dim,))
    gradient = c.input("gradient", (dim,))
    client_id = c.input("client_id")
    rnd = c.input("round")
    if derive_keys:
        key_seed = c.input("key_seed")
    else:
        master_key = c.input("master_key")
        shared_keys = c.input("shared_keys", (num_peers,))
    peer_lo = c.input("peer_lo", (num_peers,))
    peer_hi = c.input("peer_hi", (num_peers,))
    if derive_keys:
        master_key = c.poseidon([client_id, key_seed])
        shared_keys = [c.poseidon([peer_lo[j], peer_hi[j], key_seed]) for j in range(num_peers)]
    level = [T.vector_hash(c, list(features[i]) + [labels[i]]) for i in range(n)]           # computeDatasetCommitment :315-320
    tree = [level]
    while len(level) > 1:                                                                  # buildMerkleTree :198-223
        level = [c.poseidon([level[i], level[i + 1]]) for i in range(0, len(level), 2)]
        tree.append(level)
    root_w = T.vector_hash(c, weights)                                                     # weightCommitment :168-170
    root_g = T.gradient_commitment(c, gradient, client_id, rnd)                            # gradientCommitment :159-164
    root_k = c.poseidon([master_key] + list(shared_keys))                                  # keyMaterialCommitment :174-177
    masks = [[c.poseidon([shared_keys[j], rnd, peer_lo[j], peer_hi[j], k_const]) for k_const in
              [c.lin(k) for k in range(dim)]] for j in range(num_peers)]                   # derivePairwiseMask :181-196
    wires = {"tree": [[x.single_wire() for x in lvl] for lvl in tree], "root_W": root_w.single_wire(),
             "root_G": root_g.single_wire(), "root_K": root_k.single_wire(),
             "masks": [[m.single_wire() for m in row] for row in masks]}
    if derive_keys:
        wires["master"] = master_key.single_wire()
        wires["keys"] = [k.single_wire() for k in shared_keys]
    return c.compile(), wires


def compute(prover, clients: list[dict], n: int, dim: int, depth: int) -> list[dict]:
    """clients: dicts with features, labels, weights, gradient (ints, may be negative), client_id, round, peer_ids and either
    master_key + shared_keys or key_seed (keys derived on the GPU). Returns per client: tree (list of levels), root_D,
    root_W, root_G, root_K, masks (+ master, keys when derived)."""
    derive = "master_key" not in clients[0]
    prog, wires = commitment_program(n, dim, depth, len(clients[0]["peer_ids"]), derive)
    circ = prover.load_circuit(prog, check_constraints=False)
    packed = []
    for cl in clients:
        cid = cl["client_id"]
        obj = {"features": cl["features"], "labels": cl["labels"], "weights": cl["weights"], "gradient": cl["gradient"],
               "client_id": cid, "round": cl["round"],
               "peer_lo": [min(cid, j) for j in cl["peer_ids"]], "peer_hi": [max(cid, j) for j in cl["peer_ids"]]}
        if derive:
            obj["key_seed"] = cl["key_seed"]
        else:
            obj["master_key"], obj["shared_keys"] = cl["master_key"], cl["shared_keys"]
        packed.append(b"".join((int(v) % FR).to_bytes(32, "little") for v in prog.flatten_input(obj)))
    flat = [w for lvl in wires["tree"] for w in lvl] + [wires["root_W"], wires["root_G"], wires["root_K"]] + \
           [w for row in wires["masks"] for w in row]
    if derive:
        flat += [wires["master"]] + wires["keys"]
    vals = prover.eval_wires(circ, b"".join(packed), flat)
    circ.close()
    out = []
    for v in vals:
        pos, tree = 0, []
        for lvl in wires["tree"]:
            tree.append(v[pos:pos + len(lvl)])
            pos += len(lvl)
        res = {"tree": tree, "root_D": tree[-1][0], "root_W": v[pos], "root_G": v[pos + 1], "root_K": v[pos + 2]}
        pos += 3
        res["masks"] = [v[pos + j * dim:pos + (j + 1) * dim] for j in range(len(wires["masks"]))]
        pos += dim * len(wires["masks"])
        if derive:
            res["master"], res["keys"] = v[pos], v[pos + 1:pos + 1 + len(wires["keys"])]
        out.append(res)
    return out


def hydrate(prover, clients, weights, peers_of=None):
    """Fills the commitments of inputs.SimClient objects created with hashed=False -- Merkle tree and root_D, root_W, root_G,
    the key material (masterKey, K_ij, root_K) and the pairwise masks -- for ALL clients in one batched GPU pass, replacing
    the per-client circomlibjs evaluations of the reference (tests/full_system_simulation.mjs:308-335, 434-438, 567-609,
    1321-1336).  weights: one model for all clients or a list per client; peers_of(client) -> peer ids (default: the other two of
    its federation of three)."""
    if not clients:
        return clients
    if peers_of is None:
        def peers_of(c):
            base = 3 * ((c.id - 1) // 3)
            return [base + k for k in (1, 2, 3) if base + k != c.id]
    per_client = isinstance(weights[0], (list, tuple))
    req = []
    for k, c in enumerate(clients):
        c.prepare_gradient(weights[k] if per_client else weights)
        req.append({"features": c.features, "labels": c.labels, "weights": c.weights, "gradient": c.gradient, "client_id": c.id,
                    "round": c.ROUND, "key_seed": c.KEY_SEED, "peer_ids": peers_of(c)})
    c0 = clients[0]
    for c, r in zip(clients, compute(prover, req, c0.N, c0.DIM, c0.DEPTH)):
        c.attach_tree(r["tree"])
        c.pre = {"root_W": r["root_W"], "root_G": r["root_G"], "master": r["master"], "keys": r["keys"], "root_K": r["root_K"],
                 "masks": r["masks"]}
    return clients
