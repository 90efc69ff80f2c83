import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/oracle')
import zkfl_b200
from zkfl_b200 import api, formats
import parity_cases as pc
LIB='/root/repo/build/asan/libzkfl_asan.so'
P = api.Prover(0, lib_path=LIB)
from zkfl_b200 import snarkjs as sj
sj._state["lib_path"] = LIB
print('msm'); pc.case_g1_msm_degenerate(P)
for n in (1, 33, 300): pc.case_g1_msm(P, n)
pc.case_g2_msm(P, 40)
pc.case_msm_resident(P, 1030, 1)
cc = pc.tiny_circuit()
print('prove'); zk, proofs, pubs = pc.case_prove(P, cc, pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)], python_verify=0)
print('verify'); pc.case_verify_batch(P, zk, proofs, pubs)
for env in ({"ZKFL_FIXUP_QUEUE_MIN_ROWS":"1","ZKFL_MSM_CHUNK":"4"}, {"ZKFL_G2_OPERAND_FILE":"1","ZKFL_MSM_CHUNK":"7"}, {"ZKFL_MSM_CONCURRENT":"0"}, {"ZKFL_SORT_OVERLAP":"1","ZKFL_MSM_CONCURRENT":"0"}):
    os.environ.update(env); print(env)
    pc.case_g1_msm(P, 300); pc.case_g2_msm(P, 40)
    pc.case_prove(P, cc, pc.tiny_inputs(), [(11, 22), (33, 44), (0, 0)], python_verify=0)
    for k in env: os.environ.pop(k)
print('ASAN RUN OK')
if os.environ.get("ZKFL_ASAN_ROUND", "1") == "1":
    from zkfl_b200 import simulation
    rep = simulation.run_round(P, 3, setup_seed=b"asan")
    assert rep["verified"] == {"balance": 3, "training": 3, "secagg": 3}, rep
    print('ASAN ROUND OK')
