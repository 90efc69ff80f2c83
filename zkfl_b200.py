"""Import shim: loads the hyphen-named package directory as the module `zkfl_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "verifiable-federated-training-with-zero-knowledge-proofs-zk-fl-_b200")
_spec = importlib.util.spec_from_file_location(
    "zkfl_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["zkfl_b200"] = _mod
_spec.loader.exec_module(_mod)
