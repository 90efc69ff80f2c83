"""ORACLE (test infrastructure). ctypes binding of oracle/zkfl_oracle.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. Byte conventions: field elements are 32-byte little-endian canonical,
zkey points are affine Montgomery (zkey layout), results are affine canonical.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libzkfl_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "zkfl_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def ncores() -> int:
    return len(os.sched_getaffinity(0))


def _buf(n):
    return ctypes.create_string_buffer(n)


def fe(v: int) -> bytes:
    return int(v).to_bytes(32, "little")


def fes(vals) -> bytes:
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def ints(b: bytes):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def fr_mul(a: int, b: int) -> int:
    out = _buf(32)
    lib().zo_fr_mul(fe(a), fe(b), out)
    return int.from_bytes(out.raw, "little")


def fq_mul(a: int, b: int) -> int:
    out = _buf(32)
    lib().zo_fq_mul(fe(a), fe(b), out)
    return int.from_bytes(out.raw, "little")


def g1_msm(bases: bytes, scalars: bytes, nthreads: int = 0) -> bytes:
    n = len(scalars) // 32
    assert len(bases) == 64 * n
    out = _buf(64)
    lib().zo_g1_msm(bases, scalars, ctypes.c_uint64(n), out, nthreads or ncores())
    return out.raw


def g2_msm(bases: bytes, scalars: bytes, nthreads: int = 0) -> bytes:
    n = len(scalars) // 32
    assert len(bases) == 128 * n
    out = _buf(128)
    lib().zo_g2_msm(bases, scalars, ctypes.c_uint64(n), out, nthreads or ncores())
    return out.raw


def g1_mul_gen(scalars: bytes, nthreads: int = 0) -> bytes:
    n = len(scalars) // 32
    out = _buf(64 * n)
    lib().zo_g1_mul_gen(scalars, ctypes.c_uint64(n), out, nthreads or ncores())
    return out.raw


def g2_mul_gen(scalars: bytes, nthreads: int = 0) -> bytes:
    n = len(scalars) // 32
    out = _buf(128 * n)
    lib().zo_g2_mul_gen(scalars, ctypes.c_uint64(n), out, nthreads or ncores())
    return out.raw


def zkey_info(zkey: bytes):
    out = (ctypes.c_uint32 * 3)()
    rc = lib().zo_zkey_info(zkey, ctypes.c_uint64(len(zkey)), out)
    if rc:
        raise ValueError("bad zkey")
    return {"n_vars": out[0], "n_public": out[1], "domain": out[2]}


def h_scalars(zkey: bytes, wtns: bytes) -> bytes:
    info = zkey_info(zkey)
    out = _buf(32 * info["domain"])
    rc = lib().zo_h_scalars(zkey, ctypes.c_uint64(len(zkey)), wtns, out)
    assert rc == 0
    return out.raw


def groth16_prove(zkey: bytes, wtns: bytes, r: int, s: int, nthreads: int = 0):
    """wtns: n_vars*32 canonical bytes. Returns (proof 256 B, publics n_public*32 B)."""
    info = zkey_info(zkey)
    assert len(wtns) == 32 * info["n_vars"]
    proof, pub = _buf(256), _buf(32 * info["n_public"])
    rc = lib().zo_groth16_prove(zkey, ctypes.c_uint64(len(zkey)), wtns, fe(r), fe(s), proof, pub,
                                nthreads or ncores())
    assert rc == 0
    return proof.raw, pub.raw


def groth16_prove_batch(zkey: bytes, wtns: bytes, rs: bytes, nthreads: int = 0):
    info = zkey_info(zkey)
    B = len(wtns) // (32 * info["n_vars"])
    assert len(rs) == 64 * B
    proofs, pubs = _buf(256 * B), _buf(32 * info["n_public"] * B)
    rc = lib().zo_groth16_prove_batch(zkey, ctypes.c_uint64(len(zkey)), wtns, rs, B, proofs, pubs,
                                      nthreads or ncores())
    assert rc == 0
    return proofs.raw, pubs.raw


def witness_batch(prog: bytes, inputs: bytes, n_inputs: int, n_wires: int, nthreads: int = 0) -> bytes:
    B = len(inputs) // (32 * n_inputs)
    out = _buf(32 * n_wires * B)
    rc = lib().zo_witness_batch(prog, ctypes.c_uint64(len(prog)), inputs, B, out, nthreads or ncores())
    assert rc == 0
    return out.raw


def poseidon(prog: bytes, inputs) -> int:
    out = _buf(32)
    rc = lib().zo_poseidon(prog, ctypes.c_uint64(len(prog)), fes(inputs), len(inputs), out)
    assert rc == 0, rc
    return int.from_bytes(out.raw, "little")


def ntt(vals, inverse=False):
    b = ctypes.create_string_buffer(fes(vals), 32 * len(vals))
    lib().zo_ntt(b, ctypes.c_uint64(len(vals)), 1 if inverse else 0)
    return ints(b.raw)
