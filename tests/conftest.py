import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def emul_prover():
    """Prover over the HOST EMULATION of the kernels (test double; the product never loads it)."""
    import __graft_entry__ as ge
    import zkfl_b200  # noqa: F401
    from zkfl_b200.api import Prover
    p = Prover(0, lib_path=ge.build_emul())
    yield p
    p.close()


@pytest.fixture(scope="session")
def gpu_prover():
    import __graft_entry__ as ge
    import zkfl_b200  # noqa: F401
    from zkfl_b200.api import Prover
    ge.build_cuda()
    p = Prover(0)  # raises without a CUDA device: no CPU fallback
    yield p
    p.close()
