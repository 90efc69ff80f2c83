"""The reference's `component main` lines, restated (paths relative to /root/reference/src/circuits).

`build_circuit(name)` accepts the names the reference's tests pass to circom/snarkjs
(`balance_unified`, `sgd_verified`, `secure_masked_update`, ...;
tests/full_system_simulation.mjs:370-375,476-481,639-644) and returns a CompiledCircuit.
Parametrised constructors are exposed for the scaled configurations of BASELINE.json.
"""
from __future__ import annotations

from functools import lru_cache

from . import templates as T
from .builder import CircuitBuilder, CompiledCircuit, LC


def balance_proof_unified(n: int, depth: int, model_dim: int, name="balance_unified") -> CompiledCircuit:
    """balance/balance_unified.circom:74-180 (main (8,3,4) :188);
    balance/balance_unified_prod.circom:101 is the same template at (128,7,16)."""
    c = CircuitBuilder(name)
    client_id = c.input("client_id", public=True)   # noqa: F841  (bound through the public-input rows)
    root = c.input("root", public=True)
    n_public = c.input("N_public", public=True)
    c0 = c.input("c0", public=True)
    c1 = c.input("c1", public=True)
    features = c.input("features", (n, model_dim))
    labels = c.input("labels", (n,))
    siblings = c.input("siblings", (n, depth))
    path_indices = c.input("pathIndices", (n, depth))
    total = LC()
    for lab in labels:
        c.enforce(lab, lab - 1, 0)
        total = total + lab
    c.assert_eq(total, c1)
    c.assert_eq(c0 + c1, n_public)
    c.assert_eq(n_public, n)
    leaves = [T.vector_hash(c, list(features[i]) + [labels[i]]) for i in range(n)]
    T.batch_merkle_proof_prehashed(c, leaves, siblings, path_indices, root)
    return c.compile()


def training_step_verified(batch: int, model_dim: int, depth: int, precision: int,
                           name="sgd_verified") -> CompiledCircuit:
    """training/sgd_verified.circom:230-313, main (8,4,3,1000) :316."""
    c = CircuitBuilder(name)
    client_id = c.input("client_id", public=True)
    rnd = c.input("round", public=True)
    root_d = c.input("root_D", public=True)
    root_g = c.input("root_G", public=True)
    root_w = c.input("root_W", public=True)
    tau_squared = c.input("tauSquared", public=True)
    weights = c.input("weights", (model_dim,))
    expected_sum = c.input("expectedSummedGrad", (model_dim,))
    remainder = c.input("remainder", (model_dim,))
    grad_pos = c.input("gradPos", (model_dim,))
    grad_neg = c.input("gradNeg", (model_dim,))
    features = c.input("features", (batch, model_dim))
    labels = c.input("labels", (batch,))
    siblings = c.input("siblings", (batch, depth))
    path_indices = c.input("pathIndices", (batch, depth))
    # step 1 (:251-255)
    c.assert_eq(root_w, T.vector_hash(c, weights))
    # step 2 (:258-274)
    leaves = [T.vector_hash(c, list(features[i]) + [labels[i]]) for i in range(batch)]
    T.batch_merkle_proof_prehashed(c, leaves, siblings, path_indices, root_d)
    # step 3 (:277-283)
    gradient, valid = T.verify_clipping_sound(c, grad_pos, grad_neg, tau_squared, 64)
    c.assert_eq(valid, 1)
    # step 4 (:286-299)
    T.verify_gradient_correctness(c, features, labels, weights, gradient, expected_sum, remainder, precision)
    # step 5 (:302-308)
    c.assert_eq(root_g, T.gradient_commitment(c, gradient, client_id, rnd))
    return c.compile()


def training_step_plain(batch: int, model_dim: int, depth: int, *, clip_bits: int,
                        range_checks: bool, name: str) -> CompiledCircuit:
    """training/sgd_step_quick.circom:67-123 (main (8,4,3) :126, LessThan(64), no range checks) and
    training/sgd_step_v5.circom:86-166 (main (8,16,7) :168, LessThan(128) + 2^30 / 2^60 range checks)."""
    c = CircuitBuilder(name)
    client_id = c.input("client_id", public=True)
    rnd = c.input("round", public=True)
    root_d = c.input("root_D", public=True)
    root_g = c.input("root_G", public=True)
    tau_squared = c.input("tauSquared", public=True)
    grad_pos = c.input("gradPos", (model_dim,))
    grad_neg = c.input("gradNeg", (model_dim,))
    features = c.input("features", (batch, model_dim))
    labels = c.input("labels", (batch,))
    siblings = c.input("siblings", (batch, depth))
    path_indices = c.input("pathIndices", (batch, depth))
    leaves = [T.vector_hash(c, list(features[i]) + [labels[i]]) for i in range(batch)]
    T.batch_merkle_proof_prehashed(c, leaves, siblings, path_indices, root_d)
    gradient, valid = T.verify_clipping_sound(c, grad_pos, grad_neg, tau_squared, clip_bits)
    c.assert_eq(valid, 1)
    if range_checks:  # sgd_step_v5.circom:131-151
        for j in range(model_dim):
            c.assert_eq(T.less_than(c, 64, grad_pos[j], 1 << 30), 1)
            c.assert_eq(T.less_than(c, 64, grad_neg[j], 1 << 30), 1)
        c.assert_eq(T.less_than(c, 80, tau_squared, 1 << 60), 1)
    c.assert_eq(root_g, T.gradient_commitment(c, gradient, client_id, rnd))
    return c.compile()


def secure_masked_update(dim: int, num_peers: int, name="secure_masked_update") -> CompiledCircuit:
    """secureagg/secure_masked_update.circom:231-343, main (4,2) :350-360."""
    c = CircuitBuilder(name)
    client_id = c.input("client_id", public=True)
    rnd = c.input("round", public=True)
    c.input("root_D", public=True)   # binding only (:341-342)
    root_g = c.input("root_G", public=True)
    c.input("root_W", public=True)   # binding only
    root_k = c.input("root_K", public=True)
    tau_squared = c.input("tauSquared", public=True)
    masked_update = c.input("masked_update", (dim,), public=True)
    peer_ids = c.input("peer_ids", (num_peers,), public=True)
    gradient = c.input("gradient", (dim,))
    master_key = c.input("master_key")
    shared_keys = c.input("shared_keys", (num_peers,))
    c.assert_eq(root_g, T.gradient_commitment(c, gradient, client_id, rnd))       # :258-264
    c.assert_eq(root_k, c.poseidon([master_key] + list(shared_keys)))              # :270-275 (:188-200)
    T.gradient_norm_bound(c, gradient, tau_squared)                                # :281-285
    acc = list(gradient)
    masks = [T.pairwise_mask_derivation(c, shared_keys[j], rnd, client_id, peer_ids[j], dim)
             for j in range(num_peers)]                                            # :294-300
    signs = [T.sign_determination(c, client_id, peer_ids[j]) for j in range(num_peers)]  # :303-305
    for j in range(num_peers):
        acc = T.apply_signed_mask(c, acc, masks[j], signs[j])                      # :317-329
    for k in range(dim):
        c.assert_eq(masked_update[k], acc[k])                                      # :335-337
    return c.compile()


def secure_agg_client(dim: int = 8, name="secure_agg_client") -> CompiledCircuit:
    """secureagg/secure_agg_client.circom:7-163 (MainWrapper, DIM = 8)."""
    c = CircuitBuilder(name)
    client_id = c.input("client_id", public=True)
    shared_key_hash = c.input("shared_key_hash", public=True)
    root_g = c.input("root_G", public=True)
    tau_squared = c.input("tau_squared", public=True)
    masked = [c.input(f"masked_update{i}", public=True) for i in range(dim)]
    gradient = c.input("gradient", (dim,))
    mask = c.input("mask", (dim,))
    prf_seed = c.input("prf_seed")
    # GradientBoundednessProof (:22-43)
    norm = LC()
    for g in gradient:
        norm = norm + c.mul(g, g)
    c.assert_eq(T.less_than(c, 252, norm, tau_squared + 1), 1)
    # MaskDerivationProof (:45-63), PRFDerivation (:7-20)
    c.assert_eq(shared_key_hash, c.poseidon([prf_seed]))
    for i in range(dim):
        c.assert_eq(mask[i], c.poseidon([prf_seed, client_id * dim + i]))
    # MaskingCorrectnessProof (:65-72)
    for i in range(dim):
        c.assert_eq(masked[i], gradient[i] + mask[i])
    # root_G === VectorHash(gradient) (:107-111)
    c.assert_eq(root_g, T.vector_hash(c, gradient))
    return c.compile()


_MAINS = {
    "balance_unified": lambda: balance_proof_unified(8, 3, 4, "balance_unified"),
    "balance_unified_prod": lambda: balance_proof_unified(128, 7, 16, "balance_unified_prod"),
    "sgd_verified": lambda: training_step_verified(8, 4, 3, 1000, "sgd_verified"),
    "sgd_step_quick": lambda: training_step_plain(8, 4, 3, clip_bits=64, range_checks=False,
                                                  name="sgd_step_quick"),
    "sgd_step_v5": lambda: training_step_plain(8, 16, 7, clip_bits=128, range_checks=True,
                                               name="sgd_step_v5"),
    "secure_masked_update": lambda: secure_masked_update(4, 2, "secure_masked_update"),
    "secure_agg_client": lambda: secure_agg_client(8, "secure_agg_client"),
}

CIRCUIT_NAMES = tuple(_MAINS)


@lru_cache(maxsize=None)
def build_circuit(name: str) -> CompiledCircuit:
    try:
        return _MAINS[name]()
    except KeyError:
        raise KeyError(f"unknown circuit {name!r}; known: {', '.join(_MAINS)}") from None
