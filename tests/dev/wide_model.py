"""Limb-exact Python model of the even/odd wide product (no reduction) and the separate Montgomery reduction used for
lazy reduction; every carry is asserted to fit where the PTX puts it. Run: python tests/dev/wide_model.py"""
import random
M32 = (1 << 32) - 1
W = lambda x, n: [(x >> (32 * i)) & M32 for i in range(n)]
V = lambda l: sum(v << (32 * i) for i, v in enumerate(l))


def pair_chain(acc, k0, xs, y, cin=0):
    """(acc[k], acc[k+1]) += x*y for k = k0, k0+2, ... as one carry chain; returns the carry out."""
    c = cin
    for t, x in enumerate(xs):
        k = k0 + 2 * t
        pr = x * y
        lo = (pr & M32) + acc[k] + c
        acc[k] = lo & M32
        c = lo >> 32
        hi = (pr >> 32) + acc[k + 1] + c
        acc[k + 1] = hi & M32
        c = hi >> 32
    return c


def mul_wide(a, b):
    A, B = W(a, 8), W(b, 8)
    E, O = [0] * 17, [0] * 17          # O[k] has weight 2^(32(k+1)); one spare word each for the carries
    for i in range(8):
        if i % 2 == 0:
            c = pair_chain(E, i, [A[0], A[2], A[4], A[6]], B[i]); E[i + 8] += c; assert E[i + 8] <= M32
            c = pair_chain(O, i, [A[1], A[3], A[5], A[7]], B[i]); O[i + 8] += c; assert O[i + 8] <= M32
        else:
            c = pair_chain(E, i + 1, [A[1], A[3], A[5], A[7]], B[i]); E[i + 9] += c; assert E[i + 9] <= M32
            c = pair_chain(O, i - 1, [A[0], A[2], A[4], A[6]], B[i]); O[i + 7] += c; assert O[i + 7] <= M32
    assert E[16] == 0 and O[15] == 0 and O[16] == 0
    r = V(E[:16]) + (V(O[:15]) << 32)
    assert r < 1 << 512
    return r


def redc(T, p, inv, max_sub=2):
    """T (16 words) * 2^-256 mod p for T < 2*p*2^256 (result brought below p with at most two subtractions)."""
    t = W(T, 16)
    P = W(p, 8)
    X, Y = t[:8], [0] * 8              # value = X + Y << 32
    for i in range(8):
        m = (X[0] * inv) & M32
        c = pair_chain(Y, 0, [P[1], P[3], P[5], P[7]], m); assert c == 0
        c = pair_chain(X, 0, [P[0], P[2], P[4], P[6]], m)
        Y[7] += c; assert Y[7] <= M32
        assert X[0] == 0
        # divide by 2^32: X' = Y with X'[0] += X[1]; Y' = (X[2..7], 0, 0) with the carry rippling through; then bring in t[8+i]
        nx, ny = list(Y), X[2:8] + [0, 0]
        s = nx[0] + X[1]; nx[0] = s & M32; c = s >> 32
        for k in range(8):
            s = ny[k] + c; ny[k] = s & M32; c = s >> 32
        assert c == 0
        s = nx[7] + t[8 + i]; nx[7] = s & M32; c = s >> 32
        ny[7] += c; assert ny[7] <= M32
        X, Y = nx, ny
    r = V(X) + (V(Y) << 32)
    for _ in range(max_sub):
        if r >= p:
            r -= p
    assert r < p, "bound violated"
    return r


if __name__ == "__main__":
    rnd = random.Random(7)
    for p in (21888242871839275222246405745257275088696311157297823662689037894645226208583,
              21888242871839275222246405745257275088548364400416034343698204186575808495617):
        inv = (-pow(p, -1, 1 << 32)) % (1 << 32)
        Ri = pow(1 << 256, -1, p)
        pR = p << 256
        for it in range(4000):
            a, b, c, d = (rnd.randrange(p) for _ in range(4))
            if it < 3:
                a, b, c, d = [(0, 0, 0, 0), (p - 1, p - 1, p - 1, p - 1), (p - 1, p - 1, 0, 0)][it]
            assert mul_wide(a, b) == a * b
            assert redc(mul_wide(a, b), p, inv) == a * b * Ri % p
            # G1 lazy: a*b - c*d with one reduction (offset p*R keeps it non-negative)
            assert redc(mul_wide(a, b) + pR - mul_wide(c, d), p, inv) == (a * b - c * d) * Ri % p
            # Fq2 Karatsuba with lazy reduction: unreduced sums a+c, b+d < 2p
            aa, bb, s = mul_wide(a, b), mul_wide(c, d), mul_wide(a + c, b + d)
            assert a + c < 1 << 256 and b + d < 1 << 256
            assert redc(aa + pR - bb, p, inv) == (a * b - c * d) * Ri % p
            assert s - aa - bb >= 0 and redc(s - aa - bb, p, inv) == (a * d + c * b) * Ri % p
    print("wide product / separate reduction model OK")
