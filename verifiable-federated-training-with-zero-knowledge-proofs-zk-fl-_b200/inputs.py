"""Synthetic circuit inputs, restating the reference's JavaScript generators (host side, layer L4/L5 of
SURVEY section 1).  All hashing uses the product's Poseidon parameters (circuits/poseidon_params.py).

Cited reference code (paths relative to /root/reference):
  tests/full_system_simulation.mjs:118-126 (seeded LCG, evaluated in IEEE doubles like JavaScript),
  :139-238 (vectorHash / commitments / Merkle tree), :273-303 (dataset), :340-366 (balance input),
  :401-474,511-553 (verified-gradient input), :558-637 (secure-aggregation input), :1321-1336 (pairwise keys);
  tests/test_verified_gradient.mjs:162-315; tests/test_secure_aggregation.mjs:142-306;
  tests/test_secureagg.cjs:70-104; tests/balance_integration_test.mjs:35-47;
  scripts/generate_test_data_v5.mjs:19-24,56-225.
"""
from __future__ import annotations

import math

from .circuits.poseidon_params import FR, poseidon_hash

CHUNK_SIZE = 16


class JsLcg:
    """`seed = (seed * 1103515245 + 12345 [+ clientId*7919]) & 0x7fffffff` with the product evaluated as
    a double (inexact above 2^53) then ToInt32, as JavaScript does (full_system_simulation.mjs:118-122)."""

    def __init__(self, seed: int):
        self.seed = seed

    def random(self, client_term: int = 0) -> float:
        x = float(self.seed) * 1103515245.0 + 12345.0 + float(client_term) * 7919.0
        self.seed = int(x) & 0x7FFFFFFF
        return self.seed / 0x7FFFFFFF

    def random_int(self, lo: int, hi: int, client_term: int = 0) -> int:
        return math.floor(self.random(client_term) * (hi - lo + 1)) + lo


def vector_hash(values) -> int:
    values = [int(v) % FR for v in values]
    if len(values) <= CHUNK_SIZE:
        return poseidon_hash(values)
    return poseidon_hash([poseidon_hash(values[s:s + CHUNK_SIZE]) for s in range(0, len(values), CHUNK_SIZE)])


def gradient_commitment(gradient, client_id, rnd) -> int:
    return poseidon_hash([vector_hash(gradient), poseidon_hash([client_id, rnd])])


def build_merkle_tree(leaves, depth):
    leaves = list(leaves)
    zero = poseidon_hash([0])
    leaves += [zero] * ((1 << depth) - len(leaves))
    tree = [leaves]
    while len(tree[-1]) > 1:
        cur = tree[-1]
        tree.append([poseidon_hash([cur[i], cur[i + 1]]) for i in range(0, len(cur), 2)])
    return tree


def merkle_proof(tree, idx, depth):
    sib, path = [], []
    for level in range(depth):
        sib.append(tree[level][idx ^ 1])
        path.append(idx & 1)
        idx >>= 1
    return sib, path


def _s(x):
    return str(int(x))


def verified_gradient(features, labels, weights, precision=1000):
    """_computeVerifiedGradient (full_system_simulation.mjs:511-553): floor division, remainder >= 0."""
    batch, dim = len(features), len(weights)
    summed = [0] * dim
    for i in range(batch):
        err = sum(f * w for f, w in zip(features[i], weights)) - labels[i] * precision
        for j in range(dim):
            summed[j] += err * features[i][j]
    div = batch * precision
    grad = [s // div for s in summed]
    rem = [s - g * div for s, g in zip(summed, grad)]
    return grad, summed, rem


class SimClient:
    """One client of full_system_simulation.mjs (N=8, MODEL_DIM=4, DEPTH=3, tau^2=1e8, round 1)."""
    N, DIM, DEPTH, TAU2, PRECISION, ROUND = 8, 4, 3, 100000000, 1000, 1

    KEY_SEED = 12345     # masterKey = Poseidon(id, 12345), K_ij = Poseidon(min, max, 12345) (:1321-1336)

    def __init__(self, client_id: int, lcg: JsLcg, n: int | None = None, dim: int | None = None, depth: int | None = None,
                 hashed: bool = True):
        """hashed=False defers every Poseidon evaluation: the commitments then come from the GPU pipeline
        (commitments.hydrate fills tree / roots / keys / masks for all clients of a round in one batched pass)."""
        self.id = client_id
        self.pre = None      # commitments computed elsewhere: {"root_W", "root_G", "master", "keys", "root_K", "masks"}
        if n is not None:          # scaled configurations (BASELINE configs[4]); defaults are the reference's (8, 4, 3)
            self.N, self.DIM, self.DEPTH = n, dim, depth
        self.features, self.labels = [], []
        for i in range(self.N):                                           # :282-295
            self.features.append([lcg.random_int(0, 100, client_id * 1000 + i * 10 + j) for j in range(self.DIM)])
            self.labels.append((i + client_id) % 2)
        self.c1 = sum(self.labels)
        self.c0 = self.N - self.c1
        if hashed:
            leaves = [vector_hash(f + [l]) for f, l in zip(self.features, self.labels)]  # :315-320
            self.attach_tree(build_merkle_tree(leaves, self.DEPTH))

    def attach_tree(self, tree):
        self.tree = tree
        self.root_d = self.tree[-1][0]
        self.proofs = [merkle_proof(self.tree, i, self.DEPTH) for i in range(self.N)]

    def prepare_gradient(self, weights):
        """the clear gradient of this round (needed BEFORE the commitments can be computed anywhere)"""
        self.weights = list(weights)
        self.gradient, self._summed, self._rem = verified_gradient(self.features, self.labels, self.weights, self.PRECISION)
        return self.gradient

    def balance_input(self) -> dict:                                      # :355-365
        return {"client_id": _s(self.id), "root": _s(self.root_d), "N_public": _s(self.N), "c0": _s(self.c0),
                "c1": _s(self.c1), "features": [[_s(x) for x in r] for r in self.features],
                "labels": [_s(x) for x in self.labels],
                "siblings": [[_s(x) for x in p[0]] for p in self.proofs],
                "pathIndices": [[_s(x) for x in p[1]] for p in self.proofs]}

    def training_input(self, weights) -> dict:                            # :401-474
        grad = self.prepare_gradient(weights)
        summed, rem = self._summed, self._rem
        assert sum(g * g for g in grad) <= self.TAU2
        if self.pre is not None:
            self.root_w, self.root_g = self.pre["root_W"], self.pre["root_G"]
        else:
            self.root_w = vector_hash(self.weights)
            self.root_g = gradient_commitment(grad, self.id, self.ROUND)
        return {"client_id": _s(self.id), "round": _s(self.ROUND), "root_D": _s(self.root_d), "root_G": _s(self.root_g),
                "root_W": _s(self.root_w), "tauSquared": _s(self.TAU2), "weights": [_s(w) for w in self.weights],
                "expectedSummedGrad": [_s(x) for x in summed], "remainder": [_s(x) for x in rem],
                "gradPos": [_s(max(g, 0)) for g in grad], "gradNeg": [_s(max(-g, 0)) for g in grad],
                "features": [[_s(x) for x in r] for r in self.features], "labels": [_s(x) for x in self.labels],
                "siblings": [[_s(x) for x in p[0]] for p in self.proofs],
                "pathIndices": [[_s(x) for x in p[1]] for p in self.proofs]}

    def secagg_input(self, peer_ids) -> dict:                             # :558-637, keys :1321-1336
        if self.pre is not None:
            master, keys, root_k, masks = self.pre["master"], self.pre["keys"], self.pre["root_K"], self.pre["masks"]
        else:
            master = poseidon_hash([self.id, self.KEY_SEED])
            keys = [poseidon_hash([min(self.id, j), max(self.id, j), self.KEY_SEED]) for j in peer_ids]
            root_k = poseidon_hash([master] + keys)
            masks = [[poseidon_hash([key, self.ROUND, min(self.id, j), max(self.id, j), k]) for k in range(self.DIM)]
                     for j, key in zip(peer_ids, keys)]
        masked = [g % FR for g in self.gradient]
        for j, row in zip(peer_ids, masks):
            for k in range(self.DIM):
                masked[k] = (masked[k] + row[k]) % FR if self.id < j else (masked[k] - row[k]) % FR
        self.masked_update = masked
        return {"client_id": _s(self.id), "round": _s(self.ROUND), "root_D": _s(self.root_d), "root_G": _s(self.root_g),
                "root_W": _s(self.root_w), "root_K": _s(root_k), "tauSquared": _s(self.TAU2),
                "masked_update": [_s(x) for x in masked], "peer_ids": [_s(j) for j in peer_ids],
                "gradient": [_s(g % FR) for g in self.gradient], "master_key": _s(master),
                "shared_keys": [_s(k) for k in keys]}


def simulation_clients(n_clients: int = 3, seed: int = 12345):
    """Clients 1..n of full_system_simulation (one global LCG stream, clients generated in id order)."""
    lcg = JsLcg(seed)
    return [SimClient(i, lcg) for i in range(1, n_clients + 1)]


def sgd_verified_batch(n: int, seed: int = 12345, nonzero_weights: bool = False):
    """n synthetic `sgd_verified` inputs: the simulation's clients 1..n; with nonzero_weights the weights follow
    test_verified_gradient.mjs:230-235 (uniform in [-1000, 999]) from the same seeded stream."""
    lcg = JsLcg(seed)
    out = []
    for cid in range(1, n + 1):
        cl = SimClient(cid, lcg)
        w = [lcg.random_int(-1000, 999) for _ in range(cl.DIM)] if nonzero_weights else [0] * cl.DIM
        if nonzero_weights:
            cl.TAU2 = 1 << 62  # test_verified_gradient.mjs uses a large bound; stay below LessThan(64)'s range
        out.append(cl.training_input(w))
    return out


def scaled_training_input(batch: int, dim: int, depth: int, seed: int = 2024) -> dict:
    """TrainingStepVerified(batch, dim, depth, 1000) input for the scaled synthetic circuit:
    features in [0, 100], |weights| <= 1000 so every LessThan(64) operand stays below 2^64."""
    assert batch <= 1 << depth
    lcg = JsLcg(seed)
    cl = SimClient(1, lcg, n=batch, dim=dim, depth=depth)
    cl.TAU2 = 1 << 62
    return cl.training_input([lcg.random_int(-1000, 999) for _ in range(dim)])


def secure_agg_client_input() -> dict:
    """tests/test_secureagg.cjs:70-104 (DIM 8, clientId 1, prfSeed 1, gradient 0, tau^2 1)."""
    dim, cid, seed = 8, 1, 1
    mask = [poseidon_hash([seed, cid * dim + i]) for i in range(dim)]
    inp = {"client_id": _s(cid), "shared_key_hash": _s(poseidon_hash([seed])), "root_G": _s(poseidon_hash([0] * dim)),
           "tau_squared": "1", "gradient": ["0"] * dim, "mask": [_s(m) for m in mask], "prf_seed": _s(seed)}
    for i in range(dim):
        inp[f"masked_update{i}"] = _s(mask[i])
    return inp


def balance_integration_input() -> dict:
    """tests/balance_integration_test.mjs:35-47 dataset as a balance_unified(8,3,4) input."""
    labels = [0, 1, 1, 0, 1, 1, 1, 0]
    features = [[1000 + 100 * i + 1000 * j for j in range(4)] for i in range(8)]
    tree = build_merkle_tree([vector_hash(f + [l]) for f, l in zip(features, labels)], 3)
    proofs = [merkle_proof(tree, i, 3) for i in range(8)]
    return {"client_id": "1", "root": _s(tree[-1][0]), "N_public": "8", "c0": "3", "c1": "5",
            "features": [[_s(x) for x in r] for r in features], "labels": [_s(x) for x in labels],
            "siblings": [[_s(x) for x in p[0]] for p in proofs], "pathIndices": [[_s(x) for x in p[1]] for p in proofs]}


def v5_dataset(seed: int = 42, n: int = 128, dim: int = 16):
    """scripts/generate_test_data_v5.mjs:56-67."""
    lcg = JsLcg(seed)
    feats, labels = [], []
    for _ in range(n):
        feats.append([math.floor(lcg.random() * 1000) for _ in range(dim)])
        labels.append(1 if lcg.random() > 0.5 else 0)
    return feats, labels


def balance_prod_input(seed: int = 42) -> dict:
    """balance_unified_prod(128,7,16) over the seed-42 dataset behind data/test_input_v5.json (BASELINE config 1)."""
    feats, labels = v5_dataset(seed)
    tree = build_merkle_tree([vector_hash(f + [l]) for f, l in zip(feats, labels)], 7)
    proofs = [merkle_proof(tree, i, 7) for i in range(128)]
    c1 = sum(labels)
    return {"client_id": "1", "root": _s(tree[-1][0]), "N_public": "128", "c0": _s(128 - c1), "c1": _s(c1),
            "features": [[_s(x) for x in r] for r in feats], "labels": [_s(x) for x in labels],
            "siblings": [[_s(x) for x in p[0]] for p in proofs], "pathIndices": [[_s(x) for x in p[1]] for p in proofs]}
