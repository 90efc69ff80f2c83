"""Poseidon/BN254 parameters for the circuit front-end and the GPU witness evaluator.

circomlib's `poseidon_constants.circom` (included by the reference at
src/circuits/lib/poseidon.circom:17) is not vendored in the reference tree, so the
constants are regenerated here with the Grain LFSR procedure of the Poseidon paper
(x^5 S-box, R_F = 8, R_P(t) as in circomlib).  The product keeps its own generator
(this file); the oracle has an independently written one and tests compare the two.
"""
from __future__ import annotations

from functools import lru_cache

FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FULL_ROUNDS = 8
PARTIAL_ROUNDS = {t: rp for t, rp in zip(range(2, 18),
                  (56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68))}


def _grain_stream(t: int, rp: int):
    """Yields the self-shrunk Grain bit stream for (GF(p), x^alpha, n=254, t, R_F, R_P)."""
    # 80-bit state held as an int, bit 79 = oldest (b0), bit 0 = newest (b79)
    init = 0
    for value, width in ((1, 2), (0, 4), (254, 12), (t, 12), (FULL_ROUNDS, 10), (rp, 10), ((1 << 30) - 1, 30)):
        init = (init << width) | value
    state = init
    mask = (1 << 80) - 1

    def step():
        nonlocal state
        # taps at positions 0, 13, 23, 38, 51, 62 counted from the oldest bit
        b = 0
        for pos in (0, 13, 23, 38, 51, 62):
            b ^= (state >> (79 - pos)) & 1
        state = ((state << 1) & mask) | b
        return b

    for _ in range(160):
        step()
    while True:
        keep = step()
        bit = step()
        if keep:
            yield bit


@lru_cache(maxsize=None)
def poseidon_params(t: int):
    """Returns (round_constants[(8+R_P)*t], mds[t][t]) as Python ints in [0, r)."""
    rp = PARTIAL_ROUNDS[t]
    bits = _grain_stream(t, rp)

    def draw():
        v = 0
        for _ in range(254):
            v = (v << 1) | next(bits)
        return v

    consts = []
    need = (FULL_ROUNDS + rp) * t
    while len(consts) < need:
        v = draw()
        if v < FR:
            consts.append(v)
    while True:
        xs = [draw() % FR for _ in range(t)]
        ys = [draw() % FR for _ in range(t)]
        sums = [(x + y) % FR for x in xs for y in ys]
        if len(set(xs)) == t and len(set(ys)) == t and all(sums):
            break
    mds = [[pow((xs[i] + ys[j]) % FR, -1, FR) for j in range(t)] for i in range(t)]
    return consts, mds


def is_full_round(t: int, r: int) -> bool:
    half = FULL_ROUNDS // 2
    return r < half or r >= half + PARTIAL_ROUNDS[t]


def num_sboxes(t: int) -> int:
    return FULL_ROUNDS * t + PARTIAL_ROUNDS[t]


def permute(state):
    """Reference-order Poseidon permutation on Python ints (used by host helpers)."""
    t = len(state)
    consts, mds = poseidon_params(t)
    st = [x % FR for x in state]
    for r in range(FULL_ROUNDS + PARTIAL_ROUNDS[t]):
        st = [(x + consts[r * t + i]) % FR for i, x in enumerate(st)]
        if is_full_round(t, r):
            st = [pow(x, 5, FR) for x in st]
        else:
            st[0] = pow(st[0], 5, FR)
        st = [sum(m * x for m, x in zip(row, st)) % FR for row in mds]
    return st


def poseidon_hash(inputs) -> int:
    """circomlibjs `poseidon(inputs)`: capacity element 0 first, output state[0]."""
    return permute([0] + [int(x) for x in inputs])[0]
