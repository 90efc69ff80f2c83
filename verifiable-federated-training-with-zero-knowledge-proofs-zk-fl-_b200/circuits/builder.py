"""Circuit front-end: the role `circom` plays for the reference
(`circom <name>.circom --r1cs --wasm --sym`, tests/full_system_simulation.mjs:703-706).

A template written against :class:`CircuitBuilder` yields BOTH artefacts circom emits:
  * the R1CS (A.w * B.w = C.w rows, `.r1cs` iden3 layout), and
  * a straight-line witness program (`.zkwp`, our replacement for the circom `.wasm`),
    whose ops the CUDA batched witness evaluator (csrc/witness.cu) executes in lock-step
    for many client instances.

circom semantics kept: `<==` = assign + constrain, `===` =
constrain only, `<--` hints only inside Num2Bits.  Linear `<==` assignments are kept as
symbolic linear combinations (what circom's -O2 simplification does) unless a wire is
needed.  Wire order: [1 | public inputs | private inputs | internals]; none of the
in-scope `main` components has outputs.
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field

from . import poseidon_params as pp

FR = pp.FR
ONE_WIRE = 0
NONE = 0xFFFFFFFF

OP_LIN = 1       # w[dst] = <lc a>
OP_MULADD = 2    # w[dst] = <lc a> * <lc b> + <lc c>   (c may be NONE)
OP_BITS = 3      # w[dst + i] = bit i of canonical <lc a>, i < b
OP_POSEIDON = 4  # a = t, b = offset of its t-1 input wires; writes 3*nsbox+1 wires from dst


class LC:
    """Linear combination over wires (wire 0 is the constant 1). Immutable by convention."""
    __slots__ = ("t",)

    def __init__(self, terms=None):
        self.t = terms if terms is not None else {}

    @staticmethod
    def const(v: int) -> "LC":
        v %= FR
        return LC({ONE_WIRE: v} if v else {})

    @staticmethod
    def wire(w: int) -> "LC":
        return LC({w: 1})

    def _lift(o) -> "LC":  # noqa: N805
        return o if isinstance(o, LC) else LC.const(int(o))

    def __add__(self, o):
        o = LC._lift(o)
        t = dict(self.t)
        for w, c in o.t.items():
            v = (t.get(w, 0) + c) % FR
            if v:
                t[w] = v
            else:
                t.pop(w, None)
        return LC(t)

    __radd__ = __add__

    def __neg__(self):
        return LC({w: FR - c for w, c in self.t.items()})

    def __sub__(self, o):
        return self + (-LC._lift(o))

    def __rsub__(self, o):
        return LC._lift(o) + (-self)

    def __mul__(self, k):
        if isinstance(k, LC):
            raise TypeError("LC * LC is quadratic: use CircuitBuilder.mul()")
        k = int(k) % FR
        if k == 0:
            return LC()
        return LC({w: c * k % FR for w, c in self.t.items()})

    __rmul__ = __mul__

    def single_wire(self):
        if len(self.t) == 1:
            (w, c), = self.t.items()
            if c == 1 and w != ONE_WIRE:
                return w
        return None

    def eval(self, w) -> int:
        return sum(c * w[i] for i, c in self.t.items()) % FR


@dataclass
class _Matrix:
    rows: list = field(default_factory=list)
    wires: list = field(default_factory=list)
    coefs: list = field(default_factory=list)

    def add_lc(self, row: int, lc: LC):
        for w in sorted(lc.t):
            self.rows.append(row)
            self.wires.append(w)
            self.coefs.append(lc.t[w])


@dataclass
class InputSpec:
    name: str
    shape: tuple
    public: bool
    wire: int  # first wire

    @property
    def size(self):
        n = 1
        for d in self.shape:
            n *= d
        return n


def _poseidon_template(t: int):
    """Symbolic permutation for width t. Local wires: 0 = one, 1..t-1 = inputs,
    then per S-box (x^2, x^4, x^5) in round-major order, then the output wire.
    Returns (n_internal, A, B, C) with local COO matrices (row = local constraint idx)."""
    cached = _poseidon_template.cache.get(t)
    if cached:
        return cached
    consts, mds = pp.poseidon_params(t)
    rounds = pp.FULL_ROUNDS + pp.PARTIAL_ROUNDS[t]
    A, B, C = _Matrix(), _Matrix(), _Matrix()
    nxt = t  # next local wire
    row = 0
    state = [LC()] + [LC.wire(i) for i in range(1, t)]
    for r in range(rounds):
        state = [s + consts[r * t + i] for i, s in enumerate(state)]
        lanes = range(t) if pp.is_full_round(t, r) else (0,)
        for i in lanes:
            x = state[i]
            x2, x4, x5 = LC.wire(nxt), LC.wire(nxt + 1), LC.wire(nxt + 2)
            nxt += 3
            for a, b, c in ((x, x, x2), (x2, x2, x4), (x4, x, x5)):
                A.add_lc(row, a)
                B.add_lc(row, b)
                C.add_lc(row, c)
                row += 1
            state[i] = x5
        new = []
        for i in range(t):
            acc = {}
            for j in range(t):
                m = mds[i][j]
                for w, c in state[j].t.items():
                    acc[w] = (acc.get(w, 0) + m * c) % FR
            new.append(LC({w: c for w, c in acc.items() if c}))
        state = new
    out = LC.wire(nxt)
    nxt += 1
    A.add_lc(row, state[0])
    B.add_lc(row, LC.const(1))
    C.add_lc(row, out)
    row += 1
    res = (nxt - t, row, A, B, C)
    _poseidon_template.cache[t] = res
    return res


_poseidon_template.cache = {}


class CircuitBuilder:
    def __init__(self, name: str):
        self.name = name
        self.n_wires = 1
        self.inputs: list[InputSpec] = []
        self._inputs_closed = False
        self.A, self.B, self.C = _Matrix(), _Matrix(), _Matrix()
        self.n_constraints = 0
        self.ops: list[tuple] = []
        self.lcs: list[LC] = []
        self.pos_in: list[int] = []
        self.widths: set[int] = set()

    # ------------------------------------------------------------------ signals
    def input(self, name: str, shape=(), public: bool = False):
        assert not self._inputs_closed, "declare all inputs before building constraints"
        if public:
            assert all(i.public for i in self.inputs), "public inputs must be declared first"
        spec = InputSpec(name, tuple(shape), public, self.n_wires)
        self.inputs.append(spec)
        self.n_wires += spec.size

        def nest(base, dims):
            if not dims:
                return LC.wire(base), 1
            out, used = [], 0
            for _ in range(dims[0]):
                sub, n = nest(base + used, dims[1:])
                out.append(sub)
                used += n
            return out, used

        return nest(spec.wire, spec.shape)[0]

    def _new_wire(self) -> int:
        self._inputs_closed = True
        w = self.n_wires
        self.n_wires += 1
        return w

    def _lc_id(self, lc: LC) -> int:
        self.lcs.append(lc)
        return len(self.lcs) - 1

    # ------------------------------------------------------------------ constraints
    def enforce(self, a, b, c):
        """a * b === c"""
        self._inputs_closed = True
        row = self.n_constraints
        self.A.add_lc(row, LC._lift(a))
        self.B.add_lc(row, LC._lift(b))
        self.C.add_lc(row, LC._lift(c))
        self.n_constraints += 1

    def assert_eq(self, a, b):
        """a === b for linear a, b."""
        self.enforce(LC._lift(a) - LC._lift(b), LC.const(1), LC())

    def mul(self, a, b, add=None) -> LC:
        """s <== a*b (+ add): one new wire, one quadratic constraint."""
        a, b = LC._lift(a), LC._lift(b)
        w = self._new_wire()
        s = LC.wire(w)
        if add is None:
            self.ops.append((OP_MULADD, w, self._lc_id(a), self._lc_id(b), NONE))
            self.enforce(a, b, s)
        else:
            add = LC._lift(add)
            self.ops.append((OP_MULADD, w, self._lc_id(a), self._lc_id(b), self._lc_id(add)))
            self.enforce(a, b, s - add)
        return s

    def lin(self, lc) -> LC:
        """s <== lc, forcing a wire (a linear constraint)."""
        lc = LC._lift(lc)
        w = self._new_wire()
        s = LC.wire(w)
        self.ops.append((OP_LIN, w, self._lc_id(lc), 0, 0))
        self.enforce(lc, LC.const(1), s)
        return s

    def as_wire(self, lc) -> int:
        lc = LC._lift(lc)
        w = lc.single_wire()
        if w is None:
            w = self.lin(lc).single_wire()
        return w

    def num2bits(self, x, n: int):
        """circomlib Num2Bits(n): out[i] <-- (in >> i) & 1; out[i]*(out[i]-1) === 0; sum === in."""
        x = LC._lift(x)
        base = self.n_wires
        self._inputs_closed = True
        self.n_wires += n
        self.ops.append((OP_BITS, base, self._lc_id(x), n, 0))
        bits = [LC.wire(base + i) for i in range(n)]
        acc = {}
        for i, b in enumerate(bits):
            self.enforce(b, b - 1, LC())
            acc[base + i] = pow(2, i, FR)
        self.enforce(LC(acc) - x, LC.const(1), LC())
        return bits

    def poseidon(self, inputs) -> LC:
        """circomlib Poseidon(n): returns the `out` wire."""
        t = len(inputs) + 1
        assert 2 <= t <= 17
        in_wires = [self.as_wire(x) for x in inputs]
        n_internal, n_rows, tA, tB, tC = _poseidon_template(t)
        base = self.n_wires
        self._inputs_closed = True
        self.n_wires += n_internal
        gmap = [ONE_WIRE] + in_wires + list(range(base, base + n_internal))
        row0 = self.n_constraints
        for dst, src in ((self.A, tA), (self.B, tB), (self.C, tC)):
            dst.rows.extend([r + row0 for r in src.rows])
            dst.wires.extend([gmap[l] for l in src.wires])
            dst.coefs.extend(src.coefs)
        self.n_constraints += n_rows
        self.ops.append((OP_POSEIDON, base, t, len(self.pos_in), 0))
        self.pos_in.extend(in_wires)
        self.widths.add(t)
        return LC.wire(base + n_internal - 1)

    # ------------------------------------------------------------------ output
    def _schedule(self):
        """Dependency levels of the witness ops: ops of one level are independent, so the GPU evaluator runs a
        level as one launch over (ops of the level) x (client instances). Returns (ops sorted by level, offsets)."""
        wire_level = [0] * self.n_wires
        levels = []
        for op in self.ops:
            code, dst, a, b, c = op
            if code == OP_POSEIDON:
                ins = self.pos_in[b:b + a - 1]
                n_out = 3 * pp.num_sboxes(a) + 1
            else:
                ins = list(self.lcs[a].t)
                if code == OP_MULADD:
                    ins += list(self.lcs[b].t)
                    if c != NONE:
                        ins += list(self.lcs[c].t)
                n_out = b if code == OP_BITS else 1
            lvl = 1 + max((wire_level[w] for w in ins), default=0)
            for w in range(dst, dst + n_out):
                wire_level[w] = lvl
            levels.append(lvl)
        order = sorted(range(len(self.ops)), key=lambda i: levels[i])   # stable: keeps program order inside a level
        ops = [self.ops[i] for i in order]
        offs, cur = [0], 1
        for k, i in enumerate(order):
            while levels[i] > cur:
                offs.append(k)
                cur += 1
        offs.append(len(ops))
        return ops, offs

    def compile(self) -> "CompiledCircuit":
        n_pub = sum(i.size for i in self.inputs if i.public)
        n_in = sum(i.size for i in self.inputs)
        ops, level_off = self._schedule()
        return CompiledCircuit(self.name, self.n_wires, n_pub, n_in, self.n_constraints,
                               self.A, self.B, self.C, self.inputs, ops, self.lcs,
                               self.pos_in, sorted(self.widths), level_off)


def _fr_bytes(v: int) -> bytes:
    return int(v).to_bytes(32, "little")


def _container(magic: bytes, version: int, sections: list[tuple[int, bytes]]) -> bytes:
    out = [magic, struct.pack("<II", version, len(sections))]
    for sid, payload in sections:
        out.append(struct.pack("<IQ", sid, len(payload)))
        out.append(payload)
    return b"".join(out)


@dataclass
class CompiledCircuit:
    name: str
    n_wires: int
    n_public: int
    n_inputs: int
    n_constraints: int
    A: _Matrix
    B: _Matrix
    C: _Matrix
    inputs: list
    ops: list
    lcs: list
    pos_in: list
    widths: list
    level_off: list

    # -- `.r1cs` (iden3 binfile v1)
    def r1cs_bytes(self) -> bytes:
        hdr = struct.pack("<I", 32) + _fr_bytes(FR) + struct.pack(
            "<IIIIQI", self.n_wires, 0, self.n_public, self.n_inputs - self.n_public,
            self.n_wires, self.n_constraints)
        body = bytearray()
        ptr = [0, 0, 0]
        mats = (self.A, self.B, self.C)
        for row in range(self.n_constraints):
            for k, m in enumerate(mats):
                s = ptr[k]
                e = s
                rows = m.rows
                while e < len(rows) and rows[e] == row:
                    e += 1
                body += struct.pack("<I", e - s)
                for i in range(s, e):
                    body += struct.pack("<I", m.wires[i]) + _fr_bytes(m.coefs[i])
                ptr[k] = e
        labels = b"".join(struct.pack("<Q", i) for i in range(self.n_wires))
        return _container(b"r1cs", 1, [(1, hdr), (2, bytes(body)), (3, labels)])

    def input_map(self) -> dict:
        return {
            "name": self.name, "n_wires": self.n_wires, "n_public": self.n_public,
            "n_inputs": self.n_inputs, "n_constraints": self.n_constraints,
            "inputs": [{"name": i.name, "shape": list(i.shape), "public": i.public, "wire": i.wire}
                       for i in self.inputs],
        }

    # -- `.zkwp` witness program (our `.wasm` replacement)
    def program_bytes(self) -> bytes:
        lc_off = [0]
        lc_wire, lc_coef = [], bytearray()
        for lc in self.lcs:
            for w in sorted(lc.t):
                lc_wire.append(w)
                lc_coef += _fr_bytes(lc.t[w])
            lc_off.append(len(lc_wire))
        hdr = struct.pack("<8I", self.n_wires, self.n_public, self.n_inputs, len(self.ops),
                          len(self.lcs), len(lc_wire), len(self.pos_in), len(self.widths))
        ops = b"".join(struct.pack("<5I", *op) for op in self.ops)
        pos = bytearray()
        for t in self.widths:
            consts, mds = pp.poseidon_params(t)
            pos += struct.pack("<4I", t, pp.FULL_ROUNDS + pp.PARTIAL_ROUNDS[t], pp.PARTIAL_ROUNDS[t], 0)
            for c in consts:
                pos += _fr_bytes(c)
            for rowv in mds:
                for m in rowv:
                    pos += _fr_bytes(m)
        sections = [
            (1, hdr), (2, ops),
            (3, struct.pack(f"<{len(lc_off)}I", *lc_off)),
            (4, struct.pack(f"<{len(lc_wire)}I", *lc_wire)),
            (5, bytes(lc_coef)),
            (6, struct.pack(f"<{len(self.pos_in)}I", *self.pos_in)),
            (7, bytes(pos)),
            (8, json.dumps(self.input_map()).encode()),
            (9, struct.pack(f"<{len(self.level_off)}I", *self.level_off)),
        ]
        return _container(b"zkwp", 1, sections)

    # -- input.json -> flat input vector (circom semantics: decimal strings, negatives wrap)
    def flatten_input(self, obj: dict) -> list[int]:
        out = []
        for spec in self.inputs:
            if spec.name not in obj:
                raise KeyError(f"Signal not found: {spec.name}")
            flat = []

            def walk(v, dims):
                if not dims:
                    if isinstance(v, (list, tuple)):
                        raise ValueError(f"Too many values for input signal {spec.name}")
                    flat.append(int(v) % FR)
                    return
                if not isinstance(v, (list, tuple)) or len(v) != dims[0]:
                    raise ValueError(f"Wrong dimensions for input signal {spec.name}")
                for x in v:
                    walk(x, dims[1:])

            walk(obj[spec.name], spec.shape)
            out.extend(flat)
        return out
