"""dev: single-proof (B = 1) latency of sgd_verified through the C ABI, with the per-stage profile"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import zkfl_b200
from zkfl_b200 import inputs
from zkfl_b200.api import Prover
from zkfl_b200.circuits import build_circuit
P = Prover(0)
cc = build_circuit("sgd_verified")
circ = P.load_circuit(cc, check_constraints=False)
zk = P.new_zkey(cc, b"lat")
Z = P.load_zkey(zk)
ins = inputs.sgd_verified_batch(8, nonzero_weights=True)
for B in (1, 8):
    packed = circ.pack_inputs(ins[:B]); rs = [(5, 7)] * B
    for _ in range(3): P.full_prove(circ, Z, packed, rs, check=False)
    t = time.perf_counter()
    for _ in range(10): P.full_prove(circ, Z, packed, rs, check=False)
    dt = (time.perf_counter() - t) / 10
    P.prof_enable(True); P.full_prove(circ, Z, packed, rs, check=False); prof = P.prof_read(); P.prof_enable(False)
    print(f"B={B}: {dt*1e3:.2f} ms per call ({B/dt:.1f} proofs/s)", {k: round(v['ms'], 2) for k, v in prof.items()}, flush=True)
