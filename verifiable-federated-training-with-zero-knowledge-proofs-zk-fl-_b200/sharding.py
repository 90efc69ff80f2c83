"""Multi-GPU host logic: one process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Two shapes of work:
  * independent client proofs (BASELINE configs 3, 4): proof b -> rank b mod G, bases replicated on every GPU,
    NO data-path collective; the 256-byte proofs are gathered on rank 0.
  * one large proof (config 5): every rank runs the five MSMs over its own point range, the partial sums are
    all-gathered (384 B per proof per rank) and added locally, then blinded.  The witness and the H evaluations
    are computed redundantly on every rank (cheap next to the MSMs) so that is the only exchange step.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n: int, rank: int, world: int) -> list[int]:
    """proof b -> rank (b mod world)"""
    return list(range(rank, n, world))


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def gather_bytes(chunks: list[bytes], item: int) -> list[list[bytes]] | None:
    """gathers per-rank lists of fixed-size byte strings on rank 0 (returns None elsewhere)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [torch.zeros(1, dtype=torch.int64, device=_device()) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(chunks)], dtype=torch.int64, device=_device()))
    mx = max(int(c.item()) for c in counts)
    buf = torch.zeros(max(mx, 1) * item, dtype=torch.uint8, device=_device())
    if chunks:
        buf[:len(chunks) * item] = torch.frombuffer(bytearray(b"".join(chunks)), dtype=torch.uint8).to(_device())
    out = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    res = []
    for r in range(world):
        raw = bytes(out[r].cpu().numpy().tobytes())
        res.append([raw[i * item:(i + 1) * item] for i in range(int(counts[r].item()))])
    return res


def prove_independent(prover, circuit, zkey, inputs: list, rs: list | None = None):
    """Shards `inputs` round-robin; returns (proofs, publics) in input order on rank 0, (None, None) elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = shard_indices(len(inputs), rank, world)
    proofs, pubs = ([], [])
    if mine:
        proofs, pubs = prover.full_prove(circuit, zkey, [inputs[i] for i in mine], None if rs is None else [rs[i] for i in mine])
    gp = gather_bytes(proofs, 256)
    gq = gather_bytes(pubs, 32 * zkey.n_public) if zkey.n_public else None
    if rank != 0:
        return None, None
    out_p, out_q = [None] * len(inputs), [None] * len(inputs)
    for r in range(world):
        for k, i in enumerate(shard_indices(len(inputs), r, world)):
            out_p[i] = gp[r][k]
            out_q[i] = gq[r][k] if gq else b""
    return out_p, out_q


def _shared_rs(rs, B: int):
    """the blinding scalars must be the SAME on every rank (all ranks assemble the same proof): rank 0's choice -- the caller's
    or fresh CSPRNG draws -- is broadcast."""
    import secrets
    from .formats import FR
    rank = dist.get_rank()
    if rank == 0:
        vals = rs if rs is not None else [(secrets.randbelow(FR), secrets.randbelow(FR)) for _ in range(B)]
        raw = b"".join(int(r).to_bytes(32, "little") + int(s).to_bytes(32, "little") for r, s in vals)
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(_device())
    else:
        t = torch.zeros(64 * B, dtype=torch.uint8, device=_device())
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().numpy().tobytes())
    return [(int.from_bytes(raw[64 * b:64 * b + 32], "little"), int.from_bytes(raw[64 * b + 32:64 * b + 64], "little")) for b in range(B)]


def prove_split(prover, zkey, wtns: list[bytes], rs: list | None):
    """One (or a few) large proofs with every MSM split by point range over the ranks. Every rank returns the SAME proofs
    (rs=None: rank 0 draws r, s and broadcasts them)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    B = len(wtns) if isinstance(wtns, (list, tuple)) else (wtns.numel() * wtns.element_size() if hasattr(wtns, "data_ptr") else len(wtns)) // (32 * zkey.n_vars)
    rs = _shared_rs(rs, B)
    return _exchange_and_finalize(prover, zkey, wtns, B, rs)


def _exchange_and_finalize(prover, zkey, wtns, B: int, rs):
    """partials straight into a device tensor, all-gather device-to-device (NCCL over NVLink; gloo on host memory in the CPU
    tests), finalize from the gathered device tensor: no host staging on the data path."""
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.empty(384 * B, dtype=torch.uint8, device=_device())
    prover.msm_partials(zkey, wtns, rank, world, out=mine, B=B)      # returns after the library's stream is done
    allp = torch.empty(world * 384 * B, dtype=torch.uint8, device=_device())
    dist.all_gather_into_tensor(allp, mine)                          # the one exchange step: 384 B per proof per rank
    if allp.is_cuda:
        torch.cuda.current_stream().synchronize()                    # the library reads `allp` on its own stream
    return prover.finalize(zkey, allp, B, rs, nparts=world)


def full_prove_split(prover, circuit, zkey, inputs, rs: list | None, check: bool = True):
    """`fullProve` of one (or a few) large proofs over all ranks: every rank evaluates the witness on its GPU and keeps it in HBM
    (no witness ever crosses PCIe), proves its point range, then the exchange above.  Every rank returns the same proofs."""
    B = prover.witness_resident(circuit, inputs, check)
    rs = _shared_rs(rs, B)
    return _exchange_and_finalize(prover, zkey, None, B, rs)
