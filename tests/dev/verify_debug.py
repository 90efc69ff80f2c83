"""dev: run the batch verifier on the emulation library and on the CUDA library with the same inputs and compare the
intermediate buffers (which stage differs?)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import __graft_entry__ as ge
import zkfl_b200
from zkfl_b200.api import Prover
from zkfl_b200 import formats
import parity_cases as pc
E = Prover(0, lib_path=ge.EMUL)
G = Prover(0)
cc = pc.tiny_circuit()
zk, proofs, pubs = pc.case_prove(E, cc, pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)], python_verify=0)
vk = formats.vkey_json_to_bytes(formats.export_verification_key(zk))
B = len(proofs)
res = {}
for name, P in (("emul", E), ("cuda", G)):
    ok = P.verify_batch(vk, pubs, proofs)
    bufs = {}
    for buf, size in (("v_t", B * vk["n_public"] * 128), ("v_flags", 4 * B), ("v_g1", B * 3 * 80), ("v_g2", B * 144), ("v_f", (3 * B + 1) * 384), ("v_halves", 2 * B * 384)):
        out = ctypes.create_string_buffer(size)
        P._check(P.lib.zkfl_debug_read(P.ctx, buf.encode(), out, size))
        bufs[buf] = out.raw
    res[name] = (ok, bufs)
    print(name, ok)
for buf, item in (("v_t", 128), ("v_flags", 4), ("v_f", 384), ("v_halves", 384)):
    a, b = res["emul"][1][buf], res["cuda"][1][buf]
    bad = [(k // item, (k % item) // 32) for k in range(0, len(a), 32 if item >= 32 else 4) if a[k:k + min(item, 32)] != b[k:k + min(item, 32)]]
    print(buf, "equal" if not bad else f"{len(bad)} differing (item, 32-byte word): {bad[:16]}")
