"""Template library: restatement of the reference's Circom templates against CircuitBuilder.

Each function cites the template it restates (paths relative to /root/reference/src/circuits).
circomlib templates (Poseidon, Num2Bits, LessThan, LessEqThan) are third-party, restated
from circomlib 2.0.5.
"""
from __future__ import annotations

from .builder import CircuitBuilder, LC

CHUNK_SIZE = 16  # training/vector_hash.circom:52


# --------------------------------------------------------------------------- circomlib
def less_than(c: CircuitBuilder, n: int, a, b) -> LC:
    """circomlib comparators.circom LessThan(n): Num2Bits(n+1)(a + 2^n - b); out = 1 - bit_n."""
    assert n <= 252
    bits = c.num2bits(LC._lift(a) + (1 << n) - LC._lift(b), n + 1)
    return 1 - bits[n]


def less_eq_than(c: CircuitBuilder, n: int, a, b) -> LC:
    """circomlib LessEqThan(n) = LessThan(n)(a, b + 1)."""
    return less_than(c, n, a, LC._lift(b) + 1)


# --------------------------------------------------------------------------- lib/poseidon.circom, lib/merkle.circom
def poseidon_hash_n(c: CircuitBuilder, inputs) -> LC:
    """lib/poseidon.circom:35-96 (PoseidonHash1/2/N are thin wrappers over Poseidon(n))."""
    return c.poseidon(list(inputs))


def merkle_proof_verifier(c: CircuitBuilder, leaf, siblings, path_indices, root):
    """lib/merkle.circom:34-80."""
    h = LC._lift(leaf)
    for sib, bit in zip(siblings, path_indices):
        c.enforce(bit, 1 - bit, 0)                       # :58
        left = c.mul(bit, sib - h, add=h)                # :72
        right = c.mul(bit, h - sib, add=sib)             # :73
        h = c.poseidon([left, right])                    # :75
    c.assert_eq(root, h)                                 # :79


def batch_merkle_proof_prehashed(c, leaf_hashes, siblings, path_indices, root):
    """lib/merkle.circom:200-220."""
    for leaf, sibs, path in zip(leaf_hashes, siblings, path_indices):
        merkle_proof_verifier(c, leaf, sibs, path, root)


# --------------------------------------------------------------------------- training/vector_hash.circom
def vector_hash(c: CircuitBuilder, values) -> LC:
    """training/vector_hash.circom:46-89."""
    values = list(values)
    if len(values) <= CHUNK_SIZE:
        return c.poseidon(values)
    chunk_hashes = [c.poseidon(values[s:s + CHUNK_SIZE]) for s in range(0, len(values), CHUNK_SIZE)]
    return c.poseidon(chunk_hashes)


def gradient_commitment(c: CircuitBuilder, gradient, client_id, rnd) -> LC:
    """training/vector_hash.circom:195-218."""
    grad_hash = vector_hash(c, gradient)
    meta_hash = c.poseidon([client_id, rnd])
    return c.poseidon([grad_hash, meta_hash])


# --------------------------------------------------------------------------- training/sgd_*.circom
def verify_clipping_sound(c: CircuitBuilder, grad_pos, grad_neg, tau_squared, cmp_bits: int):
    """training/sgd_verified.circom:168-209 (LessThan(64)); sgd_step_quick.circom:16-49 (64);
    sgd_step_v5.circom:37-76 (LessThan(128)). Returns (gradient[], valid)."""
    for p, n in zip(grad_pos, grad_neg):
        c.enforce(p, n, 0)
    norm = LC()
    for p, n in zip(grad_pos, grad_neg):
        norm = norm + c.mul(p, p) + c.mul(n, n)
    valid = less_than(c, cmp_bits, norm, LC._lift(tau_squared) + 1)
    gradient = [p - n for p, n in zip(grad_pos, grad_neg)]
    return gradient, valid


def verify_gradient_correctness(c, features, labels, weights, claimed, expected_sum, remainder,
                                precision: int):
    """training/sgd_verified.circom:83-154 (DotProduct :39-59, SampleGradient :62-77)."""
    batch, dim = len(features), len(weights)
    sums = [LC() for _ in range(dim)]
    for i in range(batch):
        pred = LC()
        for j in range(dim):
            pred = pred + c.mul(features[i][j], weights[j])          # :46
        err = pred - labels[i] * precision                            # :70, :110
        for j in range(dim):
            sums[j] = sums[j] + c.mul(err, features[i][j])            # :74
    divisor = batch * precision
    for j in range(dim):
        c.assert_eq(expected_sum[j], sums[j])                         # :133
        lt = less_than(c, 64, remainder[j], divisor)                  # :144-147
        c.assert_eq(lt, 1)
        c.assert_eq(expected_sum[j], claimed[j] * divisor + remainder[j])  # :150


# --------------------------------------------------------------------------- secureagg/secure_masked_update.circom
def pairwise_mask_derivation(c, shared_key, rnd, client_id, peer_id, dim):
    """secureagg/secure_masked_update.circom:55-98."""
    lt = less_than(c, 64, client_id, peer_id)
    lt_client = c.mul(lt, client_id)
    lt_peer = c.mul(lt, peer_id)
    nlt_client = c.mul(1 - lt, client_id)
    nlt_peer = c.mul(1 - lt, peer_id)
    min_id = lt_client + nlt_peer
    max_id = lt_peer + nlt_client
    return [c.poseidon([shared_key, rnd, min_id, max_id, LC.const(k)]) for k in range(dim)]


def sign_determination(c, client_id, peer_id) -> LC:
    """secureagg/secure_masked_update.circom:109-119."""
    return less_than(c, 64, client_id, peer_id)


def apply_signed_mask(c, base, mask, is_positive):
    """secureagg/secure_masked_update.circom:129-146."""
    sign = 2 * LC._lift(is_positive) - 1
    return [b + c.mul(sign, m) for b, m in zip(base, mask)]


def gradient_norm_bound(c, gradient, tau_squared):
    """secureagg/secure_masked_update.circom:156-180."""
    norm = LC()
    for g in gradient:
        norm = norm + c.mul(g, g)
    c.assert_eq(less_eq_than(c, 128, norm, tau_squared), 1)
