"""The snarkjs surface the reference relies on, over libzkfl.so.

The reference's tests shell out to the snarkjs CLI (`groth16 setup|prove|verify`, `wtns calculate`, `zkey export
verificationkey`, `r1cs info`; tests/full_system_simulation.mjs:698-776,865-868), and BASELINE.json's north star names
`snarkjs.groth16.fullProve(input, wasm, zkey) -> {proof, publicSignals}`.  This module mirrors those entry points by
name, argument meaning and result shape; `cli.py` maps the command lines onto them.  The reference is JavaScript and
Node.js is not available on this image, hence the host side is Python (INTEGRATION.md shows the N-API binding).

Differences a caller can observe:
  * `<name>.wasm` holds our compiled witness program (magic `zkwp`), produced by `circom.compile`;
  * proving runs on the GPU (no CPU fallback); `fullProve` keeps the witness in HBM between the two steps;
  * `zKey.newZKey` derives the key from OS randomness (mixed with the `.ptau` bytes and the caller's entropy) instead of a
    ceremony transcript, and `zKey.contribute` re-randomises delta like snarkjs; a reproducible key needs an explicit
    `deterministic_seed=` (tests only: whoever knows the seed can forge proofs);
  * batch variants (`fullProveBatch`) prove many clients of one circuit in lock-step.
"""
from __future__ import annotations

import ctypes
import hashlib
import json
import os

from . import _lib, formats
from .api import Circuit, Prover, Zkey
from .circuits import CIRCUIT_NAMES, build_circuit

# "lib_path": None = the package's libzkfl.so.  Only the CPU test-suite sets it (to the host-emulation test double), in code;
# no environment variable can redirect the library.
_state = {"prover": None, "circuits": {}, "zkeys": {}, "lib_path": None}


def _prover() -> Prover:
    if _state["prover"] is None:
        _state["prover"] = Prover(int(os.environ.get("ZKFL_DEVICE", os.environ.get("LOCAL_RANK", "0"))), lib_path=_state["lib_path"])
    return _state["prover"]


def _file_key(path):
    st = os.stat(path)
    return (os.path.abspath(path), st.st_mtime_ns, st.st_size)


def _circuit(wasm) -> Circuit:
    """`wasm`: path of a compiled witness program (`<name>_js/<name>.wasm` written by circom.compile), or the name of a
    circuit of the built-in library."""
    if isinstance(wasm, Circuit):
        return wasm
    if isinstance(wasm, str) and wasm in CIRCUIT_NAMES:
        key = ("name", wasm)
        if key not in _state["circuits"]:
            _state["circuits"][key] = _prover().load_circuit(wasm)
        return _state["circuits"][key]
    key = _file_key(wasm)
    if key not in _state["circuits"]:
        data = open(wasm, "rb").read()
        if data[:4] != b"zkwp":
            raise ValueError(f"{wasm}: not a zkfl witness program; compile the circuit with zkfl_b200.snarkjs.circom.compile")
        base = os.path.basename(wasm).rsplit(".", 1)[0]
        r1cs_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(wasm))), base + ".r1cs")
        r1cs = open(r1cs_path, "rb").read() if os.path.exists(r1cs_path) else None
        _state["circuits"][key] = Circuit(_prover(), data, r1cs)
    return _state["circuits"][key]


def _zkey(path) -> tuple[Zkey, bytes]:
    if isinstance(path, tuple):
        return path
    key = _file_key(path)
    if key not in _state["zkeys"]:
        data = open(path, "rb").read()
        _state["zkeys"][key] = (_prover().load_zkey(data), data)
    return _state["zkeys"][key]


def _result(proof: bytes, pubs: bytes) -> dict:
    return {"proof": formats.proof_bytes_to_json(proof), "publicSignals": formats.publics_bytes_to_json(pubs)}


class circom:
    """`circom <name>.circom --r1cs --wasm --sym -o <dir>` (tests/full_system_simulation.mjs:703-706)."""

    @staticmethod
    def compile(name: str, out_dir: str = ".") -> dict:
        base = os.path.basename(name)
        if base.endswith(".circom"):
            base = base[:-7]
        cc = build_circuit(base)
        os.makedirs(os.path.join(out_dir, f"{base}_js"), exist_ok=True)
        paths = {"r1cs": os.path.join(out_dir, f"{base}.r1cs"), "wasm": os.path.join(out_dir, f"{base}_js", f"{base}.wasm"),
                 "sym": os.path.join(out_dir, f"{base}.sym")}
        open(paths["r1cs"], "wb").write(cc.r1cs_bytes())
        open(paths["wasm"], "wb").write(cc.program_bytes())
        with open(paths["sym"], "w") as f:   # label,wire,component,name for the input signals (internals are unnamed)
            f.write("0,0,0,main.one\n")
            for spec in cc.inputs:
                for k in range(spec.size):
                    f.write(f"{spec.wire + k},{spec.wire + k},0,main.{spec.name}{'[%d]' % k if spec.size > 1 else ''}\n")
        return paths


class r1cs:
    @staticmethod
    def info(path: str) -> dict:
        from .zkey_setup import r1cs_info
        return r1cs_info(open(path, "rb").read())


class wtns:
    @staticmethod
    def calculate(input, wasm, wtns_out: str | None = None) -> bytes:
        """`snarkjs wtns calculate` / `node generate_witness.cjs` (tests/full_system_simulation.mjs:760-762).
        Raises AssertFailed when a `===` of the circuit does not hold."""
        if isinstance(input, str):
            input = json.load(open(input))
        circ = _circuit(wasm)
        w = _prover().calculate_witness(circ, [input])[0]
        data = formats.wtns_write(w)
        if wtns_out:
            open(wtns_out, "wb").write(data)
        return data


class zKey:
    @staticmethod
    def newZKey(r1cs_path: str, ptau_path: str | None, zkey_out: str, entropy: str = "", deterministic_seed: bytes | None = None) -> None:
        """`snarkjs groth16 setup r1cs ptau zkey` (tests/full_system_simulation.mjs:714-717).  The toxic waste comes from the
        OS CSPRNG (mixed with `entropy` and the ptau bytes) and is discarded; `deterministic_seed` (tests only) replaces it."""
        if deterministic_seed is not None:
            seed = hashlib.sha256(b"zkfl-setup-test" + deterministic_seed).digest()
        else:
            seed = hashlib.sha512(b"zkfl-setup" + os.urandom(64) + entropy.encode()).digest()
            if ptau_path and os.path.exists(ptau_path):
                seed = hashlib.sha512(seed + open(ptau_path, "rb").read()).digest()
        data = _prover().new_zkey(open(r1cs_path, "rb").read(), seed)
        open(zkey_out, "wb").write(data)

    @staticmethod
    def contribute(zkey_in: str, zkey_out: str, name: str = "", entropy: str = "") -> None:
        """`snarkjs zkey contribute <in> <out> --name= -e=` (:726-731): delta is re-randomised with a fresh secret (CSPRNG +
        `entropy`), the C and H sections are rescaled, and a contribution record is appended (zkey_setup.contribute)."""
        data = _prover().contribute_zkey(open(zkey_in, "rb").read(), name, entropy.encode())
        open(zkey_out, "wb").write(data)

    @staticmethod
    def exportVerificationKey(zkey_path: str) -> dict:
        return formats.export_verification_key(open(zkey_path, "rb").read())


class groth16:
    @staticmethod
    def fullProve(input, wasm, zkey, rs=None) -> dict:
        if isinstance(input, str):
            input = json.load(open(input))
        return groth16.fullProveBatch([input], wasm, zkey, None if rs is None else [rs])[0]

    @staticmethod
    def fullProveBatch(inputs, wasm, zkey, rs=None) -> list:
        """Many client instances of one circuit in one GPU pass (independent proofs, shared bases)."""
        circ = _circuit(wasm)
        z, _ = _zkey(zkey)
        # one GPU pass: the constraint check (circom aborts on a failed ===) runs on the HBM-resident witness inside it and
        # raises AssertFailed; a circuit loaded without its .r1cs cannot be checked and is refused, not silently proved
        proofs, pubs = _prover().full_prove(circ, z, inputs, rs, check=True)
        return [_result(a, b) for a, b in zip(proofs, pubs)]

    @staticmethod
    def prove(zkey, wtns_file, rs=None) -> dict:
        """`snarkjs groth16 prove zkey wtns proof.json public.json` (tests/full_system_simulation.mjs:773-775)."""
        z, _ = _zkey(zkey)
        w = formats.wtns_read(open(wtns_file, "rb").read() if isinstance(wtns_file, str) else wtns_file)
        if len(w) != 32 * z.n_vars:
            raise ValueError(f"Invalid witness length. Circuit: {z.n_vars}, witness: {len(w) // 32}")
        proofs, pubs = _prover().prove(z, [w], None if rs is None else [rs])
        return _result(proofs[0], pubs[0])

    @staticmethod
    def verifyBatch(vkey: dict, items, threads: int = 0, prover: Prover | None = None, device: bool = True) -> list:
        """items: [(publicSignals, proof), ...] -> [bool, ...] (the reference runs one `snarkjs groth16 verify` process per
        proof).  device=True: all proofs in one pass of the GPU batch verifier (zkfl_groth16_verify_batch) on `prover`'s context
        (default: the module's); items that cannot be encoded (wrong count of signals, values out of range) are False.
        device=False: the single-proof host verifier on all host cores (it releases the GIL)."""
        if isinstance(vkey, str):
            vkey = json.load(open(vkey))
        if device:
            vk = formats.vkey_json_to_bytes(vkey)
            enc, slots = [], []
            for k, (sig, proof) in enumerate(items):
                try:
                    if len(sig) != vk["n_public"]:
                        raise ValueError("public signal count")
                    enc.append((b"".join(int(x).to_bytes(32, "little") for x in sig), formats.proof_json_to_bytes(proof)))
                    slots.append(k)
                except (OverflowError, ValueError):
                    pass
            res = [False] * len(items)
            if enc:
                oks = (prover or _prover()).verify_batch(vk, [e[0] for e in enc], [e[1] for e in enc])
                for k, v in zip(slots, oks):
                    res[k] = v
            return res
        from concurrent.futures import ThreadPoolExecutor
        n = threads or len(os.sched_getaffinity(0))
        with ThreadPoolExecutor(max_workers=max(1, min(n, len(items) or 1))) as ex:
            return list(ex.map(lambda it: groth16.verify(vkey, it[0], it[1]), items))

    @staticmethod
    def verify(vkey: dict, publicSignals, proof: dict) -> bool:
        """`snarkjs groth16 verify vkey public proof` (tests/full_system_simulation.mjs:865-868)."""
        if isinstance(vkey, str):
            vkey = json.load(open(vkey))
        lib = _lib.load(_state["lib_path"])
        vk = formats.vkey_json_to_bytes(vkey)
        if len(publicSignals) != vk["n_public"]:
            return False
        try:
            pubs = b"".join(int(x).to_bytes(32, "little") for x in publicSignals)
            pb = formats.proof_json_to_bytes(proof)
        except (OverflowError, ValueError):
            return False
        ok = ctypes.c_int(0)
        _lib.check(lib.zkfl_groth16_verify(_lib.as_ptr(vk["alpha1"]), _lib.as_ptr(vk["beta2"]), _lib.as_ptr(vk["gamma2"]),
                                           _lib.as_ptr(vk["delta2"]), _lib.as_ptr(vk["ic"]), _lib.as_ptr(pubs),
                                           vk["n_public"], _lib.as_ptr(pb), ctypes.byref(ok)), lib)
        return ok.value == 1
