"""ctypes binding of libzkfl.so (include/zkfl.h). No CPU fallback: if the CUDA library is missing or
no CUDA device is usable, loading / context creation raises.

The library loaded is always the package's own libzkfl.so; no environment variable redirects it.  The CPU test-suite hands the
path of the host-emulation build of the kernels (tests/_emul/, a test double, never shipped) to `load(path)` / `Prover(lib_path=)`
explicitly."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_PATH = os.path.join(_HERE, "libzkfl.so")

SYMBOLS = [
    "zkfl_last_error", "zkfl_version", "zkfl_ctx_create", "zkfl_ctx_free", "zkfl_circuit_load",
    "zkfl_circuit_free", "zkfl_circuit_info", "zkfl_zkey_load", "zkfl_zkey_load_split", "zkfl_zkey_free", "zkfl_zkey_info",
    "zkfl_r1cs_load", "zkfl_r1cs_free", "zkfl_wtns_calculate_batch", "zkfl_r1cs_check_batch", "zkfl_wtns_eval_wires",
    "zkfl_groth16_prove_batch", "zkfl_groth16_full_prove_batch", "zkfl_full_prove_stage",
    "zkfl_full_prove_run", "zkfl_full_prove_fetch", "zkfl_g1_msm", "zkfl_g2_msm", "zkfl_msm_bases_load",
    "zkfl_msm_bases_free", "zkfl_msm_run", "zkfl_g1_mul_generator", "zkfl_g2_mul_generator",
    "zkfl_launch_count", "zkfl_prof_enable", "zkfl_prof_read", "zkfl_bench_modmul", "zkfl_bench_imad", "zkfl_bench_widemac",
    "zkfl_timer_begin", "zkfl_timer_end", "zkfl_groth16_verify", "zkfl_groth16_verify_batch", "zkfl_debug_read", "zkfl_debug_pairing_selftest", "zkfl_wtns_calculate", "zkfl_groth16_prove", "zkfl_groth16_full_prove",
    "zkfl_proof_to_json", "zkfl_public_to_json", "zkfl_g1_scale_points", "zkfl_g2_scale_points", "zkfl_groth16_setup", "zkfl_aggregate_updates", "zkfl_groth16_msm_partials", "zkfl_groth16_finalize", "zkfl_ctx_wait_other",
]

_libs = {}


class ZkflError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"zkfl error {code}: {msg}")
        self.code = code


class AssertFailed(ZkflError):
    """A circuit `===` failed (circom's "Assert Failed", non-zero exit of generate_witness)."""


def library_path() -> str:
    return DEFAULT_PATH


def load(path: str | None = None):
    path = path or library_path()
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'); "
                          "zkfl_b200 has no CPU fallback")
    lib = ctypes.CDLL(path)
    vp, cp, i, u64, sz = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_size_t
    pp = ctypes.POINTER(vp)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    sig = {
        "zkfl_last_error": (cp, []), "zkfl_version": (cp, []),
        "zkfl_ctx_create": (i, [i, pp]), "zkfl_ctx_free": (None, [vp]),
        "zkfl_circuit_load": (i, [vp, vp, sz, pp]), "zkfl_circuit_free": (None, [vp]),
        "zkfl_circuit_info": (i, [vp, u32p]),
        "zkfl_zkey_load": (i, [vp, vp, sz, pp]), "zkfl_zkey_load_split": (i, [vp, vp, sz, ctypes.c_uint32, pp]), "zkfl_zkey_free": (None, [vp]), "zkfl_zkey_info": (i, [vp, u32p]),
        "zkfl_r1cs_load": (i, [vp, vp, sz, pp]), "zkfl_r1cs_free": (None, [vp]),
        "zkfl_wtns_calculate_batch": (i, [vp, vp, vp, vp, i, vp, vp]),
        "zkfl_r1cs_check_batch": (i, [vp, vp, vp, i, vp]),
        "zkfl_wtns_eval_wires": (i, [vp, vp, vp, i, vp, ctypes.c_uint32, vp]),
        "zkfl_groth16_prove_batch": (i, [vp, vp, vp, vp, i, vp, vp]),
        "zkfl_groth16_full_prove_batch": (i, [vp, vp, vp, vp, vp, vp, i, vp, vp, vp]),
        "zkfl_full_prove_stage": (i, [vp, vp, vp, vp, vp, i]),
        "zkfl_full_prove_run": (i, [vp, vp, vp, vp, i]),
        "zkfl_full_prove_fetch": (i, [vp, i, vp, vp]),
        "zkfl_g1_msm": (i, [vp, vp, vp, sz, vp]), "zkfl_g2_msm": (i, [vp, vp, vp, sz, vp]),
        "zkfl_msm_bases_load": (i, [vp, vp, sz, i, pp]), "zkfl_msm_bases_free": (None, [vp]),
        "zkfl_msm_run": (i, [vp, vp, vp, sz, vp]),
        "zkfl_g1_mul_generator": (i, [vp, vp, sz, vp]), "zkfl_g2_mul_generator": (i, [vp, vp, sz, vp]),
        "zkfl_launch_count": (u64, []),
        "zkfl_prof_enable": (i, [vp, i]), "zkfl_prof_read": (i, [vp, vp, sz]),
        "zkfl_bench_modmul": (i, [vp, sz, ctypes.c_uint32, ctypes.POINTER(ctypes.c_float)]),
        "zkfl_bench_imad": (i, [vp, sz, ctypes.c_uint32, ctypes.POINTER(ctypes.c_float)]),
        "zkfl_bench_widemac": (i, [vp, sz, ctypes.c_uint32, ctypes.POINTER(ctypes.c_float)]),
        "zkfl_groth16_verify": (i, [vp, vp, vp, vp, vp, vp, ctypes.c_uint32, vp, ctypes.POINTER(ctypes.c_int)]),
        "zkfl_groth16_verify_batch": (i, [vp, vp, vp, vp, vp, vp, ctypes.c_uint32, vp, vp, i, vp]),
        "zkfl_debug_read": (i, [vp, ctypes.c_char_p, vp, sz]),
        "zkfl_debug_pairing_selftest": (i, []),
        "zkfl_wtns_calculate": (i, [vp, vp, vp, vp, vp, vp]),
        "zkfl_groth16_prove": (i, [vp, vp, vp, vp, vp, vp, vp]),
        "zkfl_groth16_full_prove": (i, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "zkfl_proof_to_json": (i, [vp, ctypes.c_char_p, sz]),
        "zkfl_public_to_json": (i, [vp, ctypes.c_uint32, ctypes.c_char_p, sz]),
        "zkfl_groth16_msm_partials": (i, [vp, vp, vp, i, ctypes.c_uint32, ctypes.c_uint32, vp]),
        "zkfl_groth16_finalize": (i, [vp, vp, vp, ctypes.c_uint32, vp, i, vp]),
        "zkfl_ctx_wait_other": (i, [vp, vp]),
        "zkfl_aggregate_updates": (i, [vp, vp, vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_double, vp, vp, vp, vp, vp]),
        "zkfl_groth16_setup": (i, [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, vp, vp, vp, vp, vp, ctypes.c_uint32, vp, vp, sz, vp]),
        "zkfl_g1_scale_points": (i, [vp, vp, sz, vp, vp]), "zkfl_g2_scale_points": (i, [vp, vp, sz, vp, vp]),
        "zkfl_timer_begin": (i, [vp]), "zkfl_timer_end": (i, [vp, ctypes.POINTER(ctypes.c_float)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _libs[path] = lib
    return lib


def check(rc: int, lib=None):
    if rc != 0:
        msg = (lib or load()).zkfl_last_error().decode(errors="replace")
        raise (AssertFailed if rc == -5 else ZkflError)(rc, msg)


def as_ptr(buf):
    """bytes / bytearray / ctypes buffer / int address / torch tensor -> void*"""
    if buf is None:
        return None
    if isinstance(buf, int):
        return ctypes.c_void_p(buf)
    if isinstance(buf, bytes):
        return ctypes.cast(ctypes.c_char_p(buf), ctypes.c_void_p)
    if isinstance(buf, bytearray):
        return ctypes.cast((ctypes.c_char * len(buf)).from_buffer(buf), ctypes.c_void_p)
    if hasattr(buf, "data_ptr"):
        return ctypes.c_void_p(buf.data_ptr())
    return ctypes.cast(buf, ctypes.c_void_p)
