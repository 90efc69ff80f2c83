"""N > 1 host logic on CPU: two gloo ranks, each with its own context over the host emulation of the kernels.
Covers (i) independent proofs sharded b -> rank with no collective on the data path, (ii) one proof whose MSMs are
split by point range with an all-gather of the partial sums. Both must reproduce the oracle's proof bytes."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as ge
        import zkfl_b200  # noqa: F401
        import oracle_lib as ol
        import parity_cases as pc
        from zkfl_b200 import sharding
        from zkfl_b200.api import Prover
        P = Prover(0, lib_path=ge.EMUL)
        cc = pc.tiny_circuit()
        circ = P.load_circuit(cc)
        zk = open(os.path.join(ROOT, "tests", "golden", "tiny.zkey"), "rb").read()
        Z = P.load_zkey(zk, nparts=world)      # zkfl_zkey_load_split: window tables sized for a rank's share
        ins = pc.tiny_inputs()
        rs = [(11, 22), (33, 44), (55, 66)]
        assert sharding.shard_indices(5, rank, world) == ([0, 2, 4] if rank == 0 else [1, 3])
        proofs, pubs = sharding.prove_independent(P, circ, Z, ins, rs)
        ws = P.calculate_witness(circ, ins)
        ref = [ol.groth16_prove(zk, w, *r) for w, r in zip(ws, rs)]
        if rank == 0:
            assert proofs == [r[0] for r in ref] and pubs == [r[1] for r in ref]
        else:
            assert proofs is None
        split = sharding.prove_split(P, Z, ws[:2], rs[:2])
        assert split == [r[0] for r in ref[:2]], "split-MSM proof differs from the oracle"
        # fullProve split: witness evaluated and kept resident on every rank, random blinding drawn on rank 0 and broadcast
        full = sharding.full_prove_split(P, circ, Z, ins[:2], rs[:2])
        assert full == split
        rnd = sharding.full_prove_split(P, circ, Z, ins[:1], None)
        import torch
        t = torch.frombuffer(bytearray(rnd[0]), dtype=torch.uint8)
        both = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(both, t)
        assert all(bytes(x.numpy().tobytes()) == rnd[0] for x in both), "ranks returned different proofs"
        assert rnd[0] != ref[0][0]
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc() + repr(e)))
    finally:
        dist.destroy_process_group()


def _round_worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as ge
        import zkfl_b200  # noqa: F401
        from zkfl_b200 import simulation
        from zkfl_b200.api import Prover
        P = Prover(0, lib_path=ge.EMUL)
        rep = simulation.run_round(P, 3, setup_seed=b"gloo-round")       # the reference's 3-client round, proofs sharded b -> rank
        if rank == 0:
            assert rep["verified"] == {"balance": 3, "training": 3, "secagg": 3} and rep["proofs"] == 9 and rep["n_gpus"] == world, rep
            assert rep["aggregated_gradient"] == rep["expected_gradient"], rep       # masks cancel: the server recovers the mean gradient
        else:
            assert rep is None
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc() + repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_full_round():
    """simulation.run_round over two ranks (what `bench.py`'s full_round_1023 section and `torchrun -m zkfl_b200.simulation` do on
    GPUs): every phase's proofs sharded over the ranks, rank 0 verifies and aggregates, the other rank returns None, nobody waits."""
    import __graft_entry__ as ge
    ge.build_emul()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_round_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_two_rank_sharding_and_split_msm():
    import __graft_entry__ as ge
    ge.build_emul()
    ge.build_oracle()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
