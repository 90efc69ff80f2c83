"""dev check on real GPUs: torchrun --nproc-per-node G tests/dev/nccl_sharding_check.py
independent proofs sharded over ranks + one proof split by MSM point range, both against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import zkfl_b200
from zkfl_b200 import sharding, inputs as I
from zkfl_b200.api import Prover
from zkfl_b200.circuits import build_circuit
import oracle_lib as ol
P = Prover(lr)
cc = build_circuit("secure_masked_update")
circ = P.load_circuit(cc)
zk = P.new_zkey(cc, b"nccl")
Z = P.load_zkey(zk)
clients = I.simulation_clients(3)
for c in clients: c.training_input([0] * 4)
ins = [c.secagg_input([j for j in (1, 2, 3) if j != c.id]) for c in clients] * 3
rs = [(i + 1, 2 * i + 3) for i in range(len(ins))]
proofs, pubs = sharding.prove_independent(P, circ, Z, ins, rs)
ws = P.calculate_witness(circ, ins)
if rank == 0:
    ref = [ol.groth16_prove(zk, w, *r) for w, r in zip(ws, rs)]
    assert proofs == [r[0] for r in ref] and pubs == [r[1] for r in ref]
    print("independent proofs over", world, "ranks: bit-exact vs oracle")
split = sharding.prove_split(P, Z, ws[:2], rs[:2])
assert split == [ol.groth16_prove(zk, w, *r)[0] for w, r in zip(ws[:2], rs[:2])]
print(f"rank {rank}: split-MSM proof over {world} ranks bit-exact vs oracle")
dist.barrier(); dist.destroy_process_group()
