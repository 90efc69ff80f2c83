"""Host-side driver of libzkfl.so: one Prover per GPU (one process per GPU when sharding).

Batch-first: the GPU path proves B client instances of one circuit in lock-step (shared bases,
batch-minor witness layout); `B = 1` is the snarkjs single-proof case.
"""
from __future__ import annotations

import ctypes
import json
import os

from . import _lib
from .circuits import CompiledCircuit, build_circuit
from .formats import FR


def _fe_bytes(vals) -> bytes:
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


class Circuit:
    """A compiled circuit resident on the GPU (witness program + optional R1CS for `===` checks)."""

    def __init__(self, prover: "Prover", zkwp: bytes, r1cs: bytes | None = None, compiled: CompiledCircuit | None = None):
        self.prover = prover
        lib = prover.lib
        self.handle = ctypes.c_void_p()
        self._zkwp = zkwp
        prover._check(lib.zkfl_circuit_load(prover.ctx, _lib.as_ptr(zkwp), len(zkwp), ctypes.byref(self.handle)))
        info = (ctypes.c_uint32 * 4)()
        prover._check(lib.zkfl_circuit_info(self.handle, info))
        self.n_wires, self.n_public, self.n_inputs, self.n_ops = info[0], info[1], info[2], info[3]
        self.r1cs_handle = ctypes.c_void_p()
        if r1cs is not None:
            prover._check(lib.zkfl_r1cs_load(prover.ctx, _lib.as_ptr(r1cs), len(r1cs), ctypes.byref(self.r1cs_handle)))
        self.compiled = compiled
        if compiled is not None:
            self.meta = compiled.input_map()
        else:
            from .formats import read_container
            self.meta = json.loads(read_container(zkwp, b"zkwp")[8].decode())

    def flatten_input(self, obj: dict) -> list[int]:
        """input.json object -> flat signal list (circom semantics: decimal strings, negatives wrap mod r)."""
        out = []
        for spec in self.meta["inputs"]:
            name, shape = spec["name"], spec["shape"]
            if name not in obj:
                raise KeyError(f"Signal not found: {name}")
            flat = []

            def walk(v, dims):
                if not dims:
                    if isinstance(v, (list, tuple)):
                        raise ValueError(f"Too many values for input signal {name}")
                    flat.append(int(v) % FR)
                    return
                if not isinstance(v, (list, tuple)) or len(v) != dims[0]:
                    raise ValueError(f"Wrong dimensions for input signal {name}")
                for x in v:
                    walk(x, dims[1:])

            walk(obj[name], shape)
            out.extend(flat)
        return out

    def pack_inputs(self, objs) -> bytes:
        return b"".join(_fe_bytes(self.flatten_input(o)) for o in objs)

    def close(self):
        if self.handle:
            self.prover.lib.zkfl_circuit_free(self.handle)
            self.handle = ctypes.c_void_p()
        if self.r1cs_handle:
            self.prover.lib.zkfl_r1cs_free(self.r1cs_handle)
            self.r1cs_handle = ctypes.c_void_p()


class Zkey:
    def __init__(self, prover: "Prover", data: bytes, nparts: int = 1):
        """nparts > 1: key for one proof split over that many GPUs (window tables sized for a rank's share of the points)"""
        self.prover = prover
        self.handle = ctypes.c_void_p()
        if nparts > 1:
            prover._check(prover.lib.zkfl_zkey_load_split(prover.ctx, _lib.as_ptr(data), len(data), nparts, ctypes.byref(self.handle)))
        else:
            prover._check(prover.lib.zkfl_zkey_load(prover.ctx, _lib.as_ptr(data), len(data), ctypes.byref(self.handle)))
        info = (ctypes.c_uint32 * 3)()
        prover._check(prover.lib.zkfl_zkey_info(self.handle, info))
        self.n_vars, self.n_public, self.domain = info[0], info[1], info[2]

    def close(self):
        if self.handle:
            self.prover.lib.zkfl_zkey_free(self.handle)
            self.handle = ctypes.c_void_p()


class Prover:
    def __init__(self, device: int = 0, lib_path: str | None = None):
        self.lib = _lib.load(lib_path)
        self.ctx = ctypes.c_void_p()
        self._check(self.lib.zkfl_ctx_create(device, ctypes.byref(self.ctx)))
        self.device = device

    def _check(self, rc: int):
        _lib.check(rc, self.lib)

    def close(self):
        if self.ctx:
            self.lib.zkfl_ctx_free(self.ctx)
            self.ctx = ctypes.c_void_p()

    # ---------------------------------------------------------------- artefacts
    def load_circuit(self, name_or_compiled, check_constraints: bool = True) -> Circuit:
        cc = build_circuit(name_or_compiled) if isinstance(name_or_compiled, str) else name_or_compiled
        return Circuit(self, cc.program_bytes(), cc.r1cs_bytes() if check_constraints else None, cc)

    def load_circuit_files(self, zkwp_path: str, r1cs_path: str | None = None) -> Circuit:
        zkwp = open(zkwp_path, "rb").read()
        r1cs = open(r1cs_path, "rb").read() if r1cs_path and os.path.exists(r1cs_path) else None
        return Circuit(self, zkwp, r1cs)

    def load_zkey(self, data: bytes, nparts: int = 1) -> Zkey:
        return Zkey(self, data, nparts)

    def new_zkey(self, r1cs, seed: bytes | None = None) -> bytes:
        """r1cs: `.r1cs` bytes or a CompiledCircuit.  seed=None (the default, what every non-test caller should use): the toxic
        waste is drawn from the OS CSPRNG and discarded.  An explicit seed makes the key reproducible -- and FORGEABLE by anyone
        who knows the seed: tests, benchmarks and parity checks only."""
        from .zkey_setup import new_zkey
        return new_zkey(self, r1cs, os.urandom(64) if seed is None else seed)

    def contribute_zkey(self, zkey: bytes, name: str = "", entropy: bytes = b"") -> bytes:
        from .zkey_setup import contribute
        return contribute(self, zkey, name, entropy)

    def scale_points(self, pts: bytes, scalar: int, group: int = 1) -> bytes:
        """every point of a zkey section (affine Montgomery) times one scalar, on the GPU"""
        sz = 64 if group == 1 else 128
        n = len(pts) // sz
        out = ctypes.create_string_buffer(sz * n)
        fn = self.lib.zkfl_g1_scale_points if group == 1 else self.lib.zkfl_g2_scale_points
        self._check(fn(self.ctx, _lib.as_ptr(pts), n, _lib.as_ptr(int(scalar).to_bytes(32, "little")), out))
        return out.raw

    # ---------------------------------------------------------------- witness
    def calculate_witness(self, circuit: Circuit, inputs, check: bool = True) -> list[bytes]:
        """inputs: list of input.json objects (or packed bytes). Returns one canonical witness (n_wires*32 B) per client."""
        packed = inputs if isinstance(inputs, (bytes, bytearray)) else circuit.pack_inputs(inputs)
        B = len(packed) // (32 * circuit.n_inputs)
        out = ctypes.create_string_buffer(32 * circuit.n_wires * B)
        bad = (ctypes.c_uint32 * B)()
        r1 = circuit.r1cs_handle if (check and circuit.r1cs_handle) else None
        self._check(self.lib.zkfl_wtns_calculate_batch(self.ctx, circuit.handle, r1, _lib.as_ptr(packed), B, out, bad))
        sz = 32 * circuit.n_wires
        raw = memoryview(out)      # `.raw` copies the whole buffer on every access: slice one view instead
        return [bytes(raw[b * sz:(b + 1) * sz]) for b in range(B)]

    def eval_wires(self, circuit: Circuit, packed_inputs: bytes, wires: list[int]) -> list[list[int]]:
        """runs the program for every instance and returns the selected wires as Python ints, one list per instance"""
        B = len(packed_inputs) // (32 * circuit.n_inputs)
        sel = (ctypes.c_uint32 * len(wires))(*wires)
        out = ctypes.create_string_buffer(32 * len(wires) * B)
        self._check(self.lib.zkfl_wtns_eval_wires(self.ctx, circuit.handle, _lib.as_ptr(packed_inputs), B, sel, len(wires), out))
        raw, n = out.raw, len(wires)
        return [[int.from_bytes(raw[32 * (b * n + k):32 * (b * n + k + 1)], "little") for k in range(n)] for b in range(B)]

    def check_witness(self, circuit: Circuit, wtns: list[bytes]):
        B = len(wtns)
        bad = (ctypes.c_uint32 * B)()
        rc = self.lib.zkfl_r1cs_check_batch(self.ctx, circuit.r1cs_handle, _lib.as_ptr(b"".join(wtns)), B, bad)
        return rc, list(bad)

    # ---------------------------------------------------------------- prove
    @staticmethod
    def _pack_rs(rs, B):
        if rs is None:
            return None
        assert len(rs) == B
        return b"".join(int(r).to_bytes(32, "little") + int(s).to_bytes(32, "little") for r, s in rs)

    @staticmethod
    def _wtns_buffer(zkey: Zkey, wtns):
        """list of per-proof witnesses, or ONE buffer holding them back to back: bytes / bytearray, or a torch uint8 tensor
        (pinned host memory uploads at PCIe speed, a CUDA tensor is copied on the device). -> (buffer, B)"""
        if isinstance(wtns, (list, tuple)):
            return b"".join(wtns), len(wtns)
        n = wtns.numel() * wtns.element_size() if hasattr(wtns, "data_ptr") else len(wtns)
        if getattr(wtns, "is_cuda", False):
            # the library copies on its own stream: whatever produced the tensor on torch's current stream must be done first
            import torch
            torch.cuda.current_stream(wtns.device).synchronize()
        if n % (32 * zkey.n_vars):
            raise ValueError("witness buffer is not a multiple of 32 * n_vars bytes")
        return wtns, n // (32 * zkey.n_vars)

    def prove(self, zkey: Zkey, wtns, rs=None):
        """-> (list of 256-byte proofs, list of public-signal byte strings)"""
        buf, B = self._wtns_buffer(zkey, wtns)
        proofs = ctypes.create_string_buffer(256 * B)
        pubs = ctypes.create_string_buffer(max(32 * zkey.n_public * B, 1))
        self._check(self.lib.zkfl_groth16_prove_batch(self.ctx, zkey.handle, _lib.as_ptr(buf),
                                                     _lib.as_ptr(self._pack_rs(rs, B)), B, proofs, pubs))
        psz = 32 * zkey.n_public
        praw, qraw = proofs.raw, pubs.raw
        return ([praw[256 * b:256 * (b + 1)] for b in range(B)], [qraw[psz * b:psz * (b + 1)] for b in range(B)])

    def full_prove(self, circuit: Circuit, zkey: Zkey, inputs, rs=None, check: bool = True):
        """witness + prove in one GPU pass.  check=True (default): the circuit's constraints are checked on the device inside
        the pass and a failed `===` raises AssertFailed, as circom/snarkjs `fullProve` does; it needs the circuit's R1CS
        (load_circuit(..., check_constraints=True)) and raises if that is missing rather than skipping the check."""
        packed = inputs if isinstance(inputs, (bytes, bytearray)) else circuit.pack_inputs(inputs)
        B = len(packed) // (32 * circuit.n_inputs)
        proofs = ctypes.create_string_buffer(256 * B)
        pubs = ctypes.create_string_buffer(max(32 * zkey.n_public * B, 1))
        if check and not circuit.r1cs_handle:
            raise ValueError("full_prove(check=True) needs the circuit's .r1cs (it was loaded without one); pass check=False to skip the constraint check explicitly")
        bad = (ctypes.c_uint32 * B)()
        self._check(self.lib.zkfl_groth16_full_prove_batch(self.ctx, circuit.handle, zkey.handle, circuit.r1cs_handle if check else None,
                                                          _lib.as_ptr(packed), _lib.as_ptr(self._pack_rs(rs, B)), B, proofs, pubs, bad))
        psz = 32 * zkey.n_public
        praw, qraw = proofs.raw, pubs.raw
        return ([praw[256 * b:256 * (b + 1)] for b in range(B)], [qraw[psz * b:psz * (b + 1)] for b in range(B)])

    def msm_partials(self, zkey: Zkey, wtns, part: int, nparts: int, out=None, B: int | None = None):
        """the five MSM sums over this rank's point range (B x 384 bytes), see zkfl_groth16_msm_partials.
        wtns=None: the witness the last calculate_witness(..., keep_resident) left in HBM (pass B).  out: a torch uint8 tensor
        (host or CUDA) that receives the partials in place -- with a CUDA tensor nothing touches the host; default: bytes."""
        if wtns is None:
            buf = None
            assert B is not None
        else:
            buf, B = self._wtns_buffer(zkey, wtns)
        dst = out if out is not None else ctypes.create_string_buffer(384 * B)
        self._check(self.lib.zkfl_groth16_msm_partials(self.ctx, zkey.handle, _lib.as_ptr(buf), B, part, nparts,
                                                      _lib.as_ptr(dst) if out is not None else dst))
        return out if out is not None else dst.raw

    def witness_resident(self, circuit: Circuit, inputs, check: bool = True) -> int:
        """runs the witness program and LEAVES the witness in HBM (nothing is copied to the host); returns B"""
        packed = inputs if isinstance(inputs, (bytes, bytearray)) else circuit.pack_inputs(inputs)
        B = len(packed) // (32 * circuit.n_inputs)
        r1 = circuit.r1cs_handle if (check and circuit.r1cs_handle) else None
        self._check(self.lib.zkfl_wtns_calculate_batch(self.ctx, circuit.handle, r1, _lib.as_ptr(packed), B, None, None))
        return B

    def finalize(self, zkey: Zkey, partials, B: int, rs=None, nparts: int | None = None) -> list[bytes]:
        """partials: list of per-rank byte strings, or ONE torch uint8 tensor (host or CUDA) holding nparts x B x 384 bytes"""
        proofs = ctypes.create_string_buffer(256 * B)
        if isinstance(partials, (list, tuple)):
            nparts, partials = len(partials), b"".join(partials)
        self._check(self.lib.zkfl_groth16_finalize(self.ctx, zkey.handle, _lib.as_ptr(partials), nparts,
                                                  _lib.as_ptr(self._pack_rs(rs, B)), B, proofs))
        praw = proofs.raw
        return [praw[256 * b:256 * (b + 1)] for b in range(B)]

    # ---------------------------------------------------------------- MSM / setup support
    def g1_msm(self, bases: bytes, scalars: bytes) -> bytes:
        out = ctypes.create_string_buffer(64)
        self._check(self.lib.zkfl_g1_msm(self.ctx, _lib.as_ptr(bases), _lib.as_ptr(scalars), len(scalars) // 32, out))
        return out.raw

    def g2_msm(self, bases: bytes, scalars: bytes) -> bytes:
        out = ctypes.create_string_buffer(128)
        self._check(self.lib.zkfl_g2_msm(self.ctx, _lib.as_ptr(bases), _lib.as_ptr(scalars), len(scalars) // 32, out))
        return out.raw

    def g1_mul_generator(self, scalars) -> bytes:
        sc = scalars if isinstance(scalars, (bytes, bytearray)) else _fe_bytes(scalars)
        n = len(sc) // 32
        out = ctypes.create_string_buffer(64 * n)
        self._check(self.lib.zkfl_g1_mul_generator(self.ctx, _lib.as_ptr(sc), n, out))
        return out.raw

    def g2_mul_generator(self, scalars) -> bytes:
        sc = scalars if isinstance(scalars, (bytes, bytearray)) else _fe_bytes(scalars)
        n = len(sc) // 32
        out = ctypes.create_string_buffer(128 * n)
        self._check(self.lib.zkfl_g2_mul_generator(self.ctx, _lib.as_ptr(sc), n, out))
        return out.raw

    # ---------------------------------------------------------------- masked aggregation + model update on the device
    def aggregate_updates(self, masked_updates, accept, model, learning_rate: float) -> dict:
        """Server.aggregateUpdates (tests/full_system_simulation.mjs:1137-1199): masked_updates[i] = the DIM field elements client i
        published (ints), accept[i] = all of its proofs verified.  Field sum, signed decode, mean and the SGD step run on the GPU.
        -> {"aggregated_field", "aggregated_gradient" (mean), "new_model", "num_clients"}; None when no client is accepted."""
        n, dim = len(masked_updates), len(model)
        if not any(accept):
            return None
        packed = b"".join(int(v).to_bytes(32, "little") for row in masked_updates for v in row)
        acc = bytes(1 if a else 0 for a in accept)
        model_in = (ctypes.c_double * dim)(*[float(x) for x in model])
        mean, new_model = (ctypes.c_double * dim)(), (ctypes.c_double * dim)()
        field = ctypes.create_string_buffer(32 * dim)
        cnt = ctypes.c_uint32(0)
        self._check(self.lib.zkfl_aggregate_updates(self.ctx, _lib.as_ptr(packed), _lib.as_ptr(acc), n, dim, float(learning_rate), model_in,
                                                    field, mean, new_model, ctypes.byref(cnt)))
        raw = field.raw
        return {"aggregated_field": [int.from_bytes(raw[32 * j:32 * j + 32], "little") for j in range(dim)],
                "aggregated_gradient": list(mean), "new_model": list(new_model), "num_clients": cnt.value}

    # ---------------------------------------------------------------- batch verification on the device
    def verify_batch(self, vk: dict, publics: list[bytes], proofs: list[bytes]) -> list[bool]:
        """vk: formats.vkey_json_to_bytes(vkey.json); publics[b]: n_public*32 B canonical; proofs[b]: 256 B.
        One GPU pass over all proofs of a round; malformed proofs come back False."""
        B = len(proofs)
        if B == 0:
            return []
        n_public = vk["n_public"]
        ok = (ctypes.c_int32 * B)()
        self._check(self.lib.zkfl_groth16_verify_batch(self.ctx, _lib.as_ptr(vk["alpha1"]), _lib.as_ptr(vk["beta2"]),
                                                       _lib.as_ptr(vk["gamma2"]), _lib.as_ptr(vk["delta2"]), _lib.as_ptr(vk["ic"]),
                                                       n_public, _lib.as_ptr(b"".join(publics)) if n_public else None,
                                                       _lib.as_ptr(b"".join(proofs)), B, ok))
        return [v == 1 for v in ok]

    # ---------------------------------------------------------------- measurement
    def prof_enable(self, on: bool = True):
        self._check(self.lib.zkfl_prof_enable(self.ctx, 1 if on else 0))

    def prof_read(self) -> dict:
        buf = ctypes.create_string_buffer(1 << 16)
        self._check(self.lib.zkfl_prof_read(self.ctx, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, launches, calls = line.split()
            out[name] = {"ms": float(ms), "launches": int(launches), "calls": int(calls)}
        return out

    def wait_other(self, other: "Prover"):
        self._check(self.lib.zkfl_ctx_wait_other(self.ctx, other.ctx))

    def timer_begin(self):
        self._check(self.lib.zkfl_timer_begin(self.ctx))

    def timer_end(self) -> float:
        ms = ctypes.c_float()
        self._check(self.lib.zkfl_timer_end(self.ctx, ctypes.byref(ms)))
        return ms.value

    def bench_imad(self, n_threads: int, iters: int) -> float:
        """-> measured 32-bit multiply-adds per second"""
        ms = ctypes.c_float()
        self._check(self.lib.zkfl_bench_imad(self.ctx, n_threads, iters, ctypes.byref(ms)))
        return n_threads * iters * 8 / (ms.value * 1e-3)

    def bench_widemac(self, n_threads: int, iters: int) -> float:
        """-> measured fused 32x32->64 multiply-accumulates per second"""
        ms = ctypes.c_float()
        self._check(self.lib.zkfl_bench_widemac(self.ctx, n_threads, iters, ctypes.byref(ms)))
        return n_threads * iters * 4 / (ms.value * 1e-3)

    def msm_load_bases(self, bases: bytes, group: int = 1):
        h = ctypes.c_void_p()
        self._check(self.lib.zkfl_msm_bases_load(self.ctx, _lib.as_ptr(bases), len(bases) // (64 if group == 1 else 128), group, ctypes.byref(h)))
        return h

    def msm_run(self, handle, scalars, n: int, out=None):
        """scalars: host pointer/bytes (copied in) or None to reuse the scalars already staged in HBM; out None -> async"""
        self._check(self.lib.zkfl_msm_run(self.ctx, handle, _lib.as_ptr(scalars), n, _lib.as_ptr(out)))

    def msm_free_bases(self, handle):
        self.lib.zkfl_msm_bases_free(handle)

    def bench_modmul(self, n_threads: int, iters: int) -> float:
        """-> measured Montgomery products per second"""
        ms = ctypes.c_float()
        self._check(self.lib.zkfl_bench_modmul(self.ctx, n_threads, iters, ctypes.byref(ms)))
        return n_threads * iters * 2 / (ms.value * 1e-3)

    # resident (steady-state) full-prove: inputs staged once, proofs left in HBM
    def stage(self, circuit: Circuit, zkey: Zkey, packed_inputs, rs_packed, B: int):
        self._check(self.lib.zkfl_full_prove_stage(self.ctx, circuit.handle, zkey.handle, _lib.as_ptr(packed_inputs),
                                                   _lib.as_ptr(rs_packed), B))

    def run_staged(self, circuit: Circuit, zkey: Zkey, B: int, check: bool = False):
        self._check(self.lib.zkfl_full_prove_run(self.ctx, circuit.handle, zkey.handle, circuit.r1cs_handle if check else None, B))

    def fetch(self, B: int, out=None):
        buf = out if out is not None else ctypes.create_string_buffer(256 * B)
        self._check(self.lib.zkfl_full_prove_fetch(self.ctx, B, _lib.as_ptr(buf) if out is not None else buf, None))
        return buf

    def full_prove_raw(self, circuit: Circuit, zkey: Zkey, inputs_ptr, rs_ptr, B: int, proofs_ptr, pubs_ptr, check: bool = True):
        """pointer-level call (pinned host buffers) used by the end-to-end benchmark; constraint check on the device as in full_prove"""
        if check and not circuit.r1cs_handle:
            raise ValueError("full_prove_raw(check=True) needs the circuit's .r1cs")
        self._check(self.lib.zkfl_groth16_full_prove_batch(self.ctx, circuit.handle, zkey.handle, circuit.r1cs_handle if check else None,
                                                          _lib.as_ptr(inputs_ptr), _lib.as_ptr(rs_ptr), B, _lib.as_ptr(proofs_ptr),
                                                          _lib.as_ptr(pubs_ptr), None))

    def launch_count(self) -> int:
        return int(self.lib.zkfl_launch_count())
