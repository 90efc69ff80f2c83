"""ORACLE (test infrastructure, never shipped, never imported by the product).

Pure-Python big-int restatement of the BN254 arithmetic that the reference's
tool-chain (snarkjs ^0.7.5 / ffjavascript ^0.2.63 / circomlibjs ^0.1.7 /
circomlib ^2.0.5 -- un-vendored npm dependencies, /root/reference/package.json:37-46)
runs underneath every call site of the hot path
(/root/reference/tests/full_system_simulation.mjs:760-762,773-775,865-868).

Contents: Fr/Fq, Fq2, Fq12 (polynomial form), G1/G2 group law, optimal-ate pairing,
Poseidon (constants regenerated with the Grain LFSR of the Poseidon paper, the recipe
circomlib's constants come from), and the JS helper functions the reference duplicates in
every test (vectorHash / gradientCommitment / Merkle tree,
/root/reference/tests/full_system_simulation.mjs:139-238).

Parity status: Poseidon t=2,3,17 + VectorHash + Merkle + LCG are PINNED by
/root/reference/data/test_input_v5.json (see tests/test_oracle_pins.py); Poseidon
t=4,5,6 by published circomlibjs known answers; the pairing by bilinearity only.
"""
from __future__ import annotations

# ----------------------------------------------------------------------------- fields
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # base field
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # scalar field
# FIELD_PRIME at /root/reference/tests/full_system_simulation.mjs:65 is R.
MONT_BITS = 256
MONT_R_Q = (1 << MONT_BITS) % Q
MONT_R_R = (1 << MONT_BITS) % R


def inv_mod(a: int, p: int) -> int:
    return pow(a, -1, p)


def fr_root_of_unity(power: int) -> int:
    """ffjavascript convention: nqr = 5, w[28] = 5^((r-1)/2^28), w[k] = w[k+1]^2."""
    assert 0 <= power <= 28
    w = pow(5, (R - 1) >> 28, R)
    for _ in range(28 - power):
        w = w * w % R
    return w


# ----------------------------------------------------------------------------- Fq2
class Fq2:
    __slots__ = ("a", "b")  # a + b*u, u^2 = -1

    def __init__(self, a: int, b: int = 0):
        self.a = a % Q
        self.b = b % Q

    def __add__(self, o):
        return Fq2(self.a + o.a, self.b + o.b)

    def __sub__(self, o):
        return Fq2(self.a - o.a, self.b - o.b)

    def __neg__(self):
        return Fq2(-self.a, -self.b)

    def __mul__(self, o):
        if isinstance(o, int):
            return Fq2(self.a * o, self.b * o)
        return Fq2(self.a * o.a - self.b * o.b, self.a * o.b + self.b * o.a)

    __rmul__ = __mul__

    def __eq__(self, o):
        return self.a == o.a and self.b == o.b

    def is_zero(self):
        return self.a == 0 and self.b == 0

    def inv(self):
        d = inv_mod((self.a * self.a + self.b * self.b) % Q, Q)
        return Fq2(self.a * d, -self.b * d)

    def __repr__(self):
        return f"Fq2({self.a}, {self.b})"


class Fq1:
    """Thin wrapper so the generic curve code below can treat Fq like Fq2."""
    __slots__ = ("a",)

    def __init__(self, a: int):
        self.a = a % Q

    def __add__(self, o):
        return Fq1(self.a + o.a)

    def __sub__(self, o):
        return Fq1(self.a - o.a)

    def __neg__(self):
        return Fq1(-self.a)

    def __mul__(self, o):
        if isinstance(o, int):
            return Fq1(self.a * o)
        return Fq1(self.a * o.a)

    __rmul__ = __mul__

    def __eq__(self, o):
        return self.a == o.a

    def is_zero(self):
        return self.a == 0

    def inv(self):
        return Fq1(inv_mod(self.a, Q))

    def __repr__(self):
        return f"Fq1({self.a})"


# ----------------------------------------------------------------------------- curves (affine, None = infinity)
G1_GEN = (Fq1(1), Fq1(2))
G2_GEN = (
    Fq2(10857046999023057135944570762232829481370756359578518086990519993285655852781,
        11559732032986387107991004021392285783925812861821192530917403151452391805634),
    Fq2(8495653923123431417604973247489272438418190587263600148770280649306958101930,
        4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
B1 = Fq1(3)
B2 = Fq2(3, 0) * Fq2(9, 1).inv()


def is_on_curve(P, b):
    if P is None:
        return True
    x, y = P
    return y * y == x * x * x + b


def ec_neg(P):
    if P is None:
        return None
    return (P[0], -P[1])


def ec_double(P):
    if P is None:
        return None
    x, y = P
    if y.is_zero():
        return None
    lam = (x * x * 3) * (y * 2).inv()
    x3 = lam * lam - x - x
    y3 = lam * (x - x3) - y
    return (x3, y3)


def ec_add(P, S):
    if P is None:
        return S
    if S is None:
        return P
    x1, y1 = P
    x2, y2 = S
    if x1 == x2:
        if y1 == y2:
            return ec_double(P)
        return None
    lam = (y2 - y1) * (x2 - x1).inv()
    x3 = lam * lam - x1 - x2
    y3 = lam * (x1 - x3) - y1
    return (x3, y3)


def ec_mul(P, k: int):
    k %= R
    acc = None
    add = P
    while k:
        if k & 1:
            acc = ec_add(acc, add)
        add = ec_double(add)
        k >>= 1
    return acc


def ec_msm(points, scalars):
    """Plain reference MSM (double-and-add per term); small inputs only."""
    acc = None
    for P, k in zip(points, scalars):
        acc = ec_add(acc, ec_mul(P, k))
    return acc


# ----------------------------------------------------------------------------- Fq12 + pairing
# Fq12 = Fq[w] / (w^12 - 18 w^6 + 82)   (u = w^6 - 9  ->  u^2 = -1)
_FQ12_MOD = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]


class Fq12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [x % Q for x in c]

    @staticmethod
    def one():
        return Fq12([1] + [0] * 11)

    def __add__(self, o):
        return Fq12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return Fq12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return Fq12([-a for a in self.c])

    def __mul__(self, o):
        if isinstance(o, int):
            return Fq12([a * o for a in self.c])
        t = [0] * 23
        for i, a in enumerate(self.c):
            if a:
                for j, b in enumerate(o.c):
                    t[i + j] += a * b
        for i in range(22, 11, -1):
            top = t[i]
            if top:
                t[i] = 0
                t[i - 6] += 18 * top
                t[i - 12] -= 82 * top
        return Fq12(t[:12])

    __rmul__ = __mul__

    def __eq__(self, o):
        return self.c == o.c

    def is_zero(self):
        return all(a == 0 for a in self.c)

    def pow(self, e: int):
        res = Fq12.one()
        base = self
        while e:
            if e & 1:
                res = res * base
            base = base * base
            e >>= 1
        return res

    def inv(self):
        # extended Euclid over Fq[w]
        def deg(p):
            d = len(p) - 1
            while d and p[d] == 0:
                d -= 1
            return d

        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = self.c + [0], [x % Q for x in _FQ12_MOD] + [1]
        while deg(low):
            # r = high / low
            dh, dl = deg(high), deg(low)
            r = [0] * 13
            tmp = list(high)
            li = inv_mod(low[dl], Q)
            for i in range(dh - dl, -1, -1):
                r[i] = tmp[dl + i] * li % Q
                for c in range(dl + 1):
                    tmp[c + i] = (tmp[c + i] - r[i] * low[c]) % Q
            nm, new = list(hm), list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] = (nm[i + j] - lm[i] * r[j]) % Q
                    new[i + j] = (new[i + j] - low[i] * r[j]) % Q
            lm, low, hm, high = nm, new, lm, low
        li = inv_mod(low[0], Q)
        return Fq12([x * li for x in lm[:12]])


ATE_LOOP_COUNT = 29793968203157093288  # 6x+2, x = 4965661367192848881
LOG_ATE_LOOP_COUNT = 63


def _fq12_from_fq(a: int) -> Fq12:
    return Fq12([a] + [0] * 11)


_W = Fq12([0, 1] + [0] * 10)
_W2 = _W * _W
_W3 = _W2 * _W


def _twist(P):
    """G2 (over Fq2) -> curve over Fq12."""
    if P is None:
        return None
    x, y = P
    nx = Fq12([x.a - 9 * x.b, 0, 0, 0, 0, 0, x.b, 0, 0, 0, 0, 0])
    ny = Fq12([y.a - 9 * y.b, 0, 0, 0, 0, 0, y.b, 0, 0, 0, 0, 0])
    return (nx * _W2, ny * _W3)


def _cast_g1(P):
    if P is None:
        return None
    return (_fq12_from_fq(P[0].a), _fq12_from_fq(P[1].a))


def _linefunc(P1, P2, T):
    x1, y1 = P1
    x2, y2 = P2
    xt, yt = T
    if not (x1 == x2):
        m = (y2 - y1) * (x2 - x1).inv()
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1 * 3) * (y1 * 2).inv()
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def _frob12(P):
    return (P[0].pow(Q), P[1].pow(Q))


def miller_loop(Qt, Pt) -> Fq12:
    """Qt: twisted G2 point (Fq12 coords), Pt: G1 point cast to Fq12. No final exponentiation."""
    if Qt is None or Pt is None:
        return Fq12.one()
    Rp = Qt
    f = Fq12.one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f * f * _linefunc(Rp, Rp, Pt)
        Rp = ec_double(Rp)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _linefunc(Rp, Qt, Pt)
            Rp = ec_add(Rp, Qt)
    Q1 = _frob12(Qt)
    nQ2 = ec_neg(_frob12(Q1))
    f = f * _linefunc(Rp, Q1, Pt)
    Rp = ec_add(Rp, Q1)
    f = f * _linefunc(Rp, nQ2, Pt)
    return f


def final_exponentiate(f: Fq12) -> Fq12:
    return f.pow((Q ** 12 - 1) // R)


def pairing(Q2, P1) -> Fq12:
    assert is_on_curve(Q2, B2) and is_on_curve(P1, B1)
    return final_exponentiate(miller_loop(_twist(Q2), _cast_g1(P1)))


def pairing_product_is_one(pairs) -> bool:
    """pairs: [(G1 point, G2 point), ...]; checks prod e(P_i, Q_i) == 1 with one final exponentiation."""
    f = Fq12.one()
    for P1, Q2 in pairs:
        assert is_on_curve(Q2, B2) and is_on_curve(P1, B1)
        f = f * miller_loop(_twist(Q2), _cast_g1(P1))
    return final_exponentiate(f) == Fq12.one()


# ----------------------------------------------------------------------------- Poseidon
POSEIDON_RF = 8
POSEIDON_RP = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]  # t = 2..17


class _Grain:
    """Grain LFSR in self-shrinking mode (Poseidon paper, generate_parameters_grain)."""

    def __init__(self, t: int, rf: int, rp: int, n: int = 254, field: int = 1, sbox: int = 0):
        bits = []

        def push(v, w):
            for i in range(w - 1, -1, -1):
                bits.append((v >> i) & 1)

        push(field, 2)
        push(sbox, 4)
        push(n, 12)
        push(t, 12)
        push(rf, 10)
        push(rp, 10)
        bits.extend([1] * 30)
        assert len(bits) == 80
        self.s = bits
        for _ in range(160):
            self._raw()

    def _raw(self):
        s = self.s
        nb = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(nb)
        return nb

    def bit(self):
        while True:
            a = self._raw()
            b = self._raw()
            if a:
                return b

    def draw(self, n: int = 254) -> int:
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bit()
        return v


_POSEIDON_CACHE: dict[int, tuple[list[int], list[list[int]]]] = {}


def poseidon_constants(t: int):
    """(C[(RF+RP)*t], M[t][t]) for width t; identical to circomlib's poseidon_constants."""
    if t in _POSEIDON_CACHE:
        return _POSEIDON_CACHE[t]
    rp = POSEIDON_RP[t - 2]
    g = _Grain(t, POSEIDON_RF, rp)
    C = []
    while len(C) < (POSEIDON_RF + rp) * t:
        v = g.draw()
        if v < R:
            C.append(v)
    while True:
        xs = [g.draw() % R for _ in range(t)]
        ys = [g.draw() % R for _ in range(t)]
        ok = len(set(xs)) == t and len(set(ys)) == t
        ok = ok and all((x + y) % R != 0 for x in xs for y in ys)
        if ok:
            break
    M = [[inv_mod((xs[i] + ys[j]) % R, R) for j in range(t)] for i in range(t)]
    _POSEIDON_CACHE[t] = (C, M)
    return C, M


def poseidon_perm(state: list[int]) -> list[int]:
    t = len(state)
    C, M = poseidon_constants(t)
    rp = POSEIDON_RP[t - 2]
    st = list(state)
    for r in range(POSEIDON_RF + rp):
        st = [(x + C[r * t + i]) % R for i, x in enumerate(st)]
        if r < POSEIDON_RF // 2 or r >= POSEIDON_RF // 2 + rp:
            st = [pow(x, 5, R) for x in st]
        else:
            st[0] = pow(st[0], 5, R)
        st = [sum(M[i][j] * st[j] for j in range(t)) % R for i in range(t)]
    return st


def poseidon(inputs) -> int:
    """circomlibjs poseidon(inputs): state = [0, inputs...], output state[0]."""
    assert 1 <= len(inputs) <= 16
    return poseidon_perm([0] + [int(x) % R for x in inputs])[0]


# ----------------------------------------------------------------------------- reference JS helpers
CHUNK_SIZE = 16  # /root/reference/tests/full_system_simulation.mjs:63


def vector_hash(values) -> int:
    """/root/reference/tests/full_system_simulation.mjs:139-155; circuit twin
    /root/reference/src/circuits/training/vector_hash.circom:46-89."""
    values = [int(v) for v in values]
    if len(values) <= CHUNK_SIZE:
        return poseidon(values)
    chunk_hashes = []
    for s in range(0, len(values), CHUNK_SIZE):
        chunk_hashes.append(poseidon(values[s:s + CHUNK_SIZE]))
    return poseidon(chunk_hashes)


def gradient_commitment(gradient_field, client_id, rnd) -> int:
    """/root/reference/tests/full_system_simulation.mjs:159-164."""
    return poseidon([vector_hash(gradient_field), poseidon([client_id, rnd])])


def weight_commitment(weights) -> int:
    """/root/reference/tests/full_system_simulation.mjs:168-170 (negatives wrap mod r)."""
    return vector_hash([int(w) % R for w in weights])


def key_material_commitment(master_key, peer_keys) -> int:
    """/root/reference/tests/full_system_simulation.mjs:174-177."""
    return poseidon([master_key] + list(peer_keys))


def derive_pairwise_mask(shared_key, rnd, client_id, peer_id, dim):
    """/root/reference/tests/full_system_simulation.mjs:181-196."""
    lo, hi = min(client_id, peer_id), max(client_id, peer_id)
    return [poseidon([shared_key, rnd, lo, hi, k]) for k in range(dim)]


def build_merkle_tree(leaf_hashes, depth):
    """/root/reference/tests/full_system_simulation.mjs:198-223."""
    leaves = list(leaf_hashes)
    zero_hash = poseidon([0])
    while len(leaves) < (1 << depth):
        leaves.append(zero_hash)
    tree = [leaves]
    cur = leaves
    while len(cur) > 1:
        cur = [poseidon([cur[i], cur[i + 1]]) for i in range(0, len(cur), 2)]
        tree.append(cur)
    return tree


def get_merkle_proof(tree, leaf_idx, depth):
    """/root/reference/tests/full_system_simulation.mjs:225-238."""
    siblings, path = [], []
    idx = leaf_idx
    for level in range(depth):
        siblings.append(tree[level][idx ^ 1])
        path.append(idx % 2)
        idx //= 2
    return siblings, path


class JsLcg:
    """The tests' LCG evaluated like JavaScript does: the product is an IEEE double
    (inexact above 2^53), then ToInt32 and `& 0x7fffffff`.
    /root/reference/tests/full_system_simulation.mjs:118-126,
    /root/reference/scripts/generate_test_data_v5.mjs:19-24."""

    def __init__(self, seed: int):
        self.seed = seed

    def random(self, client_term: int = 0) -> float:
        x = float(self.seed) * 1103515245.0 + 12345.0 + float(client_term) * 7919.0
        self.seed = int(x) & 0x7FFFFFFF
        return self.seed / 0x7FFFFFFF

    def random_int(self, lo: int, hi: int, client_term: int = 0) -> int:
        import math
        return math.floor(self.random(client_term) * (hi - lo + 1)) + lo
