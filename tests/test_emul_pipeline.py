"""CPU suite for the kernel logic: the CUDA sources compiled as a host emulation (tests/_emul, a test
double that the product never loads) must agree with the oracle bit for bit. The GPU suite
(tests/test_gpu_parity.py) repeats these cases on the real sm_100a library at full sizes."""
import pytest

import parity_cases as pc
from zkfl_b200 import _lib
from zkfl_b200 import inputs as I
from zkfl_b200.circuits import build_circuit


def test_generator_mul(emul_prover):
    pc.case_generator_mul(emul_prover, n1=6, n2=2)


def test_g1_msm_sizes_and_edges(emul_prover):
    for n in (1, 2, 33, 300):
        pc.case_g1_msm(emul_prover, n)
    pc.case_g1_msm_degenerate(emul_prover)
    pc.case_linearity(emul_prover, n=32)


def test_g2_msm(emul_prover):
    pc.case_g2_msm(emul_prover, 40)


def test_tiny_circuit_witness_prove_verify(emul_prover):
    cc = pc.tiny_circuit()
    pc.case_witness(emul_prover, cc, pc.tiny_inputs())
    pc.case_prove(emul_prover, cc, pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)])


def test_failed_constraint_raises_assert(emul_prover):
    cc = pc.tiny_circuit()
    circ = emul_prover.load_circuit(cc)
    with pytest.raises(_lib.AssertFailed):
        emul_prover.calculate_witness(circ, [{"out": "5", "bound": "17", "x": "3", "y": "5"}])
    with pytest.raises(_lib.AssertFailed):   # x >= bound
        emul_prover.calculate_witness(circ, [{"out": str(15 ** 2 + 3), "bound": "3", "x": "3", "y": "5"}])
    circ.close()


def test_sgd_verified_witness_batch(emul_prover):
    cc = build_circuit("sgd_verified")
    pc.case_witness(emul_prover, cc, I.sgd_verified_batch(2) + I.sgd_verified_batch(1, nonzero_weights=True))


def test_bad_arguments_fail_loudly(emul_prover):
    with pytest.raises(_lib.ZkflError):
        emul_prover.load_zkey(b"zkey" + bytes(100))
    with pytest.raises(_lib.ZkflError):
        emul_prover.g1_mul_generator((2 ** 256 - 1).to_bytes(32, "little"))   # not reduced mod r
