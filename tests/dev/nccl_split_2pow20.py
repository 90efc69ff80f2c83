"""dev measurement on real GPUs (BASELINE configs[4]): ONE proof of TrainingStepVerified(256, 32, 8, 1000) -- domain 2^20 --
with every MSM split by point range over the ranks and the partial sums exchanged with an NCCL all-gather.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 tests/dev/nccl_split_2pow20.py
Rank 0 builds the circuit and the key (GPU setup) and shares the key through a file; every rank loads it, proves its share,
all-gathers, finalises.  Reported: wall-clock latency of the blocking call, max over ranks (the path stages 384 B per rank
through the host, so a device-only timer would not cover it), next to the whole proof on one GPU."""
import json, os, sys, time
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
from zkfl_b200.circuits.library import training_step_verified

P = Prover(lr)
t = time.time()
BATCH, DIM, DEPTH = (int(x) for x in os.environ.get("ZKFL_SCALED", "256,32,8").split(","))   # 512,32,9 -> 2^21; 960,32,10 -> 2^22
cc = training_step_verified(BATCH, DIM, DEPTH, 1000, f"sgd_scaled_{BATCH}_{DIM}_{DEPTH}")
inp = I.scaled_training_input(BATCH, DIM, DEPTH)
key_path = f"/tmp/zkfl_scaled_{BATCH}_{DIM}_{DEPTH}.zkey"
if rank == 0:
    zk = P.new_zkey(cc, b"scaled")
    open(key_path + ".tmp", "wb").write(zk)
    os.replace(key_path + ".tmp", key_path)
dist.barrier()
if rank != 0:
    zk = open(key_path, "rb").read()
Z = P.load_zkey(zk)
circ = P.load_circuit(cc, check_constraints=False)
ws = P.calculate_witness(circ, [inp], check=False)
if rank == 0:
    print(f"build + setup + load {time.time() - t:.0f} s, zkey {len(zk) / 1e6:.0f} MB, domain {Z.domain}", flush=True)
rs = [(3, 4)]


def timed(fn, reps=5):
    out, best = None, []
    for i in range(reps + 1):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.time()
        out = fn()
        dt = torch.tensor([time.time() - t0], device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if i:                       # first call allocates the workspace
            best.append(float(dt.item()))
    return out, sorted(best)[len(best) // 2]


whole, t_whole = timed(lambda: P.prove(Z, ws, rs)[0])
split, t_split = timed(lambda: sharding.prove_split(P, Z, ws, rs))
assert split == whole, "split proof differs from the whole proof"
# the same with the witness in pinned host memory and resident on the device (no pageable staging copy)
wpin = torch.frombuffer(bytearray(ws[0]), dtype=torch.uint8).pin_memory()
wdev = wpin.cuda()
whole_pin, t_whole_pin = timed(lambda: P.prove(Z, wpin, rs)[0])
split_pin, t_split_pin = timed(lambda: sharding.prove_split(P, Z, wpin, rs))
whole_dev, t_whole_dev = timed(lambda: P.prove(Z, wdev, rs)[0])
split_dev, t_split_dev = timed(lambda: sharding.prove_split(P, Z, wdev, rs))
assert whole_pin == whole and split_pin == whole and whole_dev == whole and split_dev == whole
P.prof_enable(True); P.prove(Z, ws, rs); prof_whole = P.prof_read()
sharding.prove_split(P, Z, ws, rs); prof_split = P.prof_read(); P.prof_enable(False)
if rank == 0:
    print("stages, whole proof (ms):", {k: round(v["ms"], 2) for k, v in prof_whole.items()}, flush=True)
    print(f"stages, rank 0's share of {world} (ms):", {k: round(v["ms"], 2) for k, v in prof_split.items()}, flush=True)
    import oracle_lib as ol
    t_cpu = None
    if not os.environ.get("ZKFL_SKIP_ORACLE"):
        t0 = time.time()
        ref_p, _ = ol.groth16_prove(zk, ws[0], 3, 4)
        t_cpu = time.time() - t0
        assert whole[0] == ref_p
    from zkfl_b200 import formats, snarkjs as sj
    assert sj.groth16.verify(formats.export_verification_key(zk), formats.publics_bytes_to_json(P.prove(Z, ws, rs)[1][0]), formats.proof_bytes_to_json(whole[0]))
    print(json.dumps({"circuit": f"TrainingStepVerified({BATCH},{DIM},{DEPTH},1000)", "n_wires": cc.n_wires, "domain": Z.domain, "n_gpus": world,
                      "whole_proof_one_gpu_ms": round(1e3 * t_whole, 2), "split_proof_ms": round(1e3 * t_split, 2),
                      "pinned_witness": {"whole_ms": round(1e3 * t_whole_pin, 2), "split_ms": round(1e3 * t_split_pin, 2)},
                      "device_witness": {"whole_ms": round(1e3 * t_whole_dev, 2), "split_ms": round(1e3 * t_split_dev, 2)},
                      "oracle_cpu_s": None if t_cpu is None else round(t_cpu, 2), "oracle_threads": ol.ncores(),
                      "bit_exact_vs_oracle": t_cpu is not None, "verified": True}), flush=True)
dist.barrier(); dist.destroy_process_group()
