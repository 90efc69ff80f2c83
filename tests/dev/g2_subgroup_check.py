"""Numerical check (oracle arithmetic) of the endomorphism-based G2 subgroup test used by the RLC batch verifier:
[x+1]P + psi([x]P) + psi^2([x]P) == psi^3([2x]P)  <=>  P in the order-r subgroup of the twist (BN254)."""
import os, sys, random
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "oracle"))
import bn254_ref as bn

X = 4965661367192848881
XI = bn.Fq2(9, 1)


def fq2_pow(a, e):
    r = bn.Fq2(1, 0)
    while e:
        if e & 1:
            r = r * a
        a = a * a
        e >>= 1
    return r


G12 = fq2_pow(XI, (bn.Q - 1) // 3)
G13 = fq2_pow(XI, (bn.Q - 1) // 2)


def conj(a):
    return bn.Fq2(a.a, -a.b)


def psi(P):
    if P is None:
        return None
    return (conj(P[0]) * G12, conj(P[1]) * G13)


def mul_raw(P, k):   # bn.ec_mul reduces k mod r, which is wrong off the subgroup
    acc = None
    while k:
        if k & 1:
            acc = bn.ec_add(acc, P)
        P = bn.ec_double(P)
        k >>= 1
    return acc


def in_subgroup(P):
    a = mul_raw(P, X)
    b = psi(a)
    a = bn.ec_add(a, P)
    res = psi(b)
    c = bn.ec_add(bn.ec_add(res, b), a)
    res = psi(res)
    res = bn.ec_double(res)
    return bn.ec_add(res, bn.ec_neg(c)) is None


def random_twist_point(rng):
    while True:
        x = bn.Fq2(rng.randrange(bn.Q), rng.randrange(bn.Q))
        y2 = x * x * x + bn.B2
        # sqrt in Fq2 by exponent (q^2 + 7)/16 style is messy: use y2^((q^2+1)/4)?  q^2 = 1 mod 8 -> try Tonelli via brute candidates
        y = fq2_sqrt(y2)
        if y is not None:
            return (x, y)


def fq2_sqrt(a):
    # q = 3 mod 4: algorithm 9 of "Square root computation over even extension fields" (Adj, Rodriguez-Henriquez)
    q = bn.Q
    a1 = fq2_pow(a, (q - 3) // 4)
    alpha = a1 * a1 * a
    a0 = fq2_pow(alpha, q) * alpha
    if a0 == bn.Fq2(q - 1, 0):
        return None
    x0 = a1 * a
    if alpha == bn.Fq2(q - 1, 0):
        return bn.Fq2(0, 1) * x0
    b = fq2_pow(alpha + bn.Fq2(1, 0), (q - 1) // 2)
    return b * x0


if __name__ == "__main__":
    rng = random.Random(5)
    for _ in range(3):
        k = rng.randrange(1, bn.R)
        assert in_subgroup(bn.ec_mul(bn.G2_GEN, k))
    bad = 0
    for _ in range(3):
        P = random_twist_point(rng)
        assert bn.is_on_curve(P, bn.B2)
        assert not (mul_raw(P, bn.R) is None), "random twist point landed in the subgroup?"
        bad += 0 if in_subgroup(P) else 1
    assert bad == 3, bad
    print("subgroup check formula ok")
