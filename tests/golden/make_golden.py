"""Regenerates tests/golden/. Run in the build container (reads /root/reference; the GPU box has no reference tree).
  * test_input_v5.json : verbatim copy of /root/reference/data/test_input_v5.json (the reference's only fixture)
  * poseidon_kats.json : circomlibjs known answers quoted in SURVEY.md section 8c item 5
  * tiny.zkey / tiny.proof : pure-Python oracle setup + proof of tests/parity_cases.tiny_circuit (r=11, s=22)
"""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

shutil.copy("/root/reference/data/test_input_v5.json", os.path.join(HERE, "test_input_v5.json"))
kats = [
    ([1, 2], "7853200120776062878684798364095072458815029376092732009249414926327459813530"),
    ([1, 2, 3, 4], "18821383157269793795438455681495246036402687001665670618754263018637548127333"),
    (list(range(1, 17)), "9989051620750914585850546081941653841776809718687451684622678807385399211877"),
    ([1, 2, 3], "6542985608222806190361240322586112750744169038454362455181422643027100751666"),
    ([1, 2, 3, 4, 5], "6183221330272524995739186171720101788151706631170188140075976616310159254464"),
    ([0], "19014214495641488759237505126948346942972912379615652741039992445865937985820"),
    ([1], "18586133768512220936620570745912940619677854269274689475585506675881198879027"),
]
json.dump([{"inputs": i, "hash": h} for i, h in kats], open(os.path.join(HERE, "poseidon_kats.json"), "w"), indent=1)

import groth16_ref as g16  # noqa: E402
import witness_ref as wr  # noqa: E402
import zkfl_b200  # noqa: E402,F401
from parity_cases import tiny_circuit, tiny_inputs  # noqa: E402

cc = tiny_circuit()
r1 = wr.R1cs(cc.r1cs_bytes())
zk, _ = g16.setup(r1, tau=123456789, alpha=987654321, beta=55555, delta=7777777)
open(os.path.join(HERE, "tiny.zkey"), "wb").write(zk)
w = wr.calculate_witness(wr.Program(cc.program_bytes()), cc.flatten_input(tiny_inputs()[0]))
proof, _ = g16.prove(zk, w, r=11, s=22)
open(os.path.join(HERE, "tiny.proof"), "wb").write(g16.proof_to_bytes(proof))
print("golden written")
