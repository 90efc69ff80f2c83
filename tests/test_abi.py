"""The C-ABI library loads and exports every symbol include/zkfl.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zkfl.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(zkfl_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    import zkfl_b200  # noqa: F401
    from zkfl_b200 import _lib
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build_cuda()   # no-op when libzkfl.so is newer than its sources
    out = subprocess.check_output(["nm", "-D", "--defined-only", ge.LIB], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    lib = ctypes.CDLL(ge.LIB)           # loads without a GPU (cudart is linked statically)
    lib.zkfl_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.zkfl_version()


def test_product_fails_loudly_without_a_device():
    """no CPU fallback: on a box without a GPU the product context cannot be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import zkfl_b200  # noqa: F401
    from zkfl_b200 import _lib
    from zkfl_b200.api import Prover
    with pytest.raises(_lib.ZkflError):
        Prover(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "verifiable-federated-training-with-zero-knowledge-proofs-zk-fl-_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in src and "bn254_ref" not in src and "groth16_ref" not in src, f


def test_host_verifier_in_the_real_library_without_a_gpu():
    """`groth16 verify` is host code inside libzkfl.so (CPU): it must accept the golden proof made by the
    pure-Python oracle and reject tampered inputs, with no CUDA device present."""
    import __graft_entry__ as ge
    ge.build_cuda()
    import zkfl_b200  # noqa: F401
    from zkfl_b200 import formats
    from zkfl_b200 import snarkjs as sj
    g = os.path.join(ROOT, "tests", "golden")
    zk = open(os.path.join(g, "tiny.zkey"), "rb").read()
    proof = formats.proof_bytes_to_json(open(os.path.join(g, "tiny.proof"), "rb").read())
    vk = formats.export_verification_key(zk)
    assert sj.groth16.verify(vk, ["228", "17"], proof)
    assert not sj.groth16.verify(vk, ["229", "17"], proof)
    assert not sj.groth16.verify(vk, ["228"], proof)
    assert not sj.groth16.verify(vk, [str(2 ** 255), "17"], proof)
    bad = dict(proof, pi_a=[proof["pi_a"][1], proof["pi_a"][0], "1"])
    assert not sj.groth16.verify(vk, ["228", "17"], bad)
    # the two views of Fq12, the Frobenius maps and both Miller loops agree in the nvcc host build as well
    from zkfl_b200 import _lib
    assert _lib.load().zkfl_debug_pairing_selftest() == 0


def test_napi_addon_source_compiles_and_binds_only_declared_symbols():
    """bindings/node/zkfl_napi.cc (the in-process Node.js binding of INTEGRATION.md section 2) cannot be built into an addon here
    (no Node.js headers on the image), but its source must stay in step with the C ABI: it compiles against the declaration stub
    of <node_api.h> with include/zkfl.h's real prototypes (so every call has the right arity and types), and every zkfl_* it
    calls is a symbol the header declares."""
    src = os.path.join(ROOT, "bindings", "node", "zkfl_napi.cc")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-DNODE_GYP_MODULE_NAME=zkfl_napi",
                           "-I" + os.path.join(ROOT, "tests", "stubs"), "-I" + os.path.join(ROOT, "include"), src])
    code = re.sub(r"//[^\n]*", "", open(src).read())
    used = set(re.findall(r"\b(zkfl_[a-z0-9_]+)\s*\(", code))
    assert used and used <= set(declared_symbols()), used - set(declared_symbols())
    assert {"zkfl_groth16_full_prove_batch", "zkfl_groth16_verify_batch", "zkfl_wtns_calculate_batch", "zkfl_groth16_prove_batch"} <= used
