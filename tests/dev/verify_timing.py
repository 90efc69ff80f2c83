"""dev: batch verifier timing on one GPU -- wall clock of the blocking call and device stage times, repeated calls"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import zkfl_b200
from zkfl_b200.api import Prover
from zkfl_b200 import formats, inputs as I, snarkjs as sj
from zkfl_b200.circuits import build_circuit
P = Prover(0)
cc = build_circuit("sgd_verified")
circ = P.load_circuit(cc)
zk = P.new_zkey(cc, b"vt")
Z = P.load_zkey(zk)
ins = I.sgd_verified_batch(4)
proofs, pubs = P.full_prove(circ, Z, ins, [(1, 2), (3, 4), (5, 6), (7, 8)])
vkj = formats.export_verification_key(zk)
vk = formats.vkey_json_to_bytes(vkj)
modes = {"rlc+coop (default)": {}, "coop per proof": {"ZKFL_VERIFY_RLC": "0"}, "thread per proof (round 1)": {"ZKFL_VERIFY_COOP": "0"}}
for B, mode in [(b, m) for m in modes for b in (1, 32, 256, 1024, 3072, 8192)]:
    for k in ("ZKFL_VERIFY_RLC", "ZKFL_VERIFY_COOP"):
        os.environ.pop(k, None)
    os.environ.update(modes[mode])
    ps = [proofs[i % 4] for i in range(B)]; qs = [pubs[i % 4] for i in range(B)]
    for rep in range(3):
        P.prof_enable(True)
        t = time.perf_counter(); ok = P.verify_batch(vk, qs, ps); dt = time.perf_counter() - t
        prof = P.prof_read(); P.prof_enable(False)
        assert all(ok)
    print(f"[{mode}] B={B}: {dt * 1e3:.1f} ms wall ({B / dt:.0f} proofs/s)", {k: round(v["ms"], 1) for k, v in prof.items()}, flush=True)
for k in ("ZKFL_VERIFY_RLC", "ZKFL_VERIFY_COOP"):
    os.environ.pop(k, None)
# one tampered proof in a batch of 3072: the combined check fails, the per-proof form names it
ps = [proofs[i % 4] for i in range(3072)]; qs = [pubs[i % 4] for i in range(3072)]
ps[1000] = ps[1000][:192] + ps[1000][:64]
for rep in range(2):
    t = time.perf_counter(); ok = P.verify_batch(vk, qs, ps); dt = time.perf_counter() - t
assert ok.count(False) == 1 and not ok[1000]
print(f"3072 proofs, one tampered (combined check + per-proof fallback): {dt * 1e3:.1f} ms wall", flush=True)
t = time.perf_counter()
ok = sj.groth16.verifyBatch(vkj, [(formats.publics_bytes_to_json(q), formats.proof_bytes_to_json(p)) for p, q in zip(proofs * 16, pubs * 16)], device=False)
dt = time.perf_counter() - t
print(f"host verifier, 64 proofs on {len(os.sched_getaffinity(0))} threads: {dt * 1e3:.0f} ms ({64 / dt:.0f} proofs/s)", all(ok))
