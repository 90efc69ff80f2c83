#!/usr/bin/env python
"""bench.py -- Groth16 proofs/s on the verified-gradient circuit (BASELINE.json `metric`, configs[1]).

A step = one pass of the hot path (witness -> constraint check -> H -> 5 MSMs -> blinding) over one batch of B synthetic
client instances of `sgd_verified` = TrainingStepVerified(8,4,3,1000) per GPU.  Independent proofs shard across GPUs with
no data-path collective (weak scaling: B proofs per rank).

  value : proofs/s with inputs already resident in HBM (device-event timed)
  e2e   : proofs/s through the C ABI full-prove call with pinned HOST buffers (H2D/D2H inside the timed region)
  roofline : the dominant kernel (G1 bucket accumulation) in fused 32x32->64 multiply-accumulates per second against (frac)
             the live-measured rate of the same instruction and (frac_nominal) 32 per clock per SM at the sampled SM clock;
             roofline_g2 / roofline_ntt: the G2 accumulation and the H-polynomial stage against the same integer rooflines
             (at this size the NTT is bound by its Montgomery products, not by HBM -- its GB/s is reported beside it)
  msm_g1_2pow20 : BASELINE.json's second metric, one 2^20-point G1 MSM over resident bases, with its own roofline and stages
  split_proof : BASELINE configs[4]: ONE proof of the 2^20-domain synthetic training circuit split over the N GPUs of the run
             (strong scaling; NCCL all-gather of the partial sums device-to-device)
  cpu_baseline : the C++ oracle (restatement of snarkjs' algorithm) on the box's host cores, bounded sample

`--impl reference` times the CPU arm alone: snarkjs itself cannot run here (no Node.js on the image), so it is
the C++ oracle port with all host threads (cpu_baseline.kind = "port").
`--workload split` makes the split proof the headline line (latency in ms, lower is better, strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "groth16_proofs_per_s_verified_gradient"
UNIT = "proofs/s"
WORKLOAD = "sgd_verified TrainingStepVerified(8,4,3,1000): witness+prove, batch of B client proofs per GPU per step"
SPLIT_SHAPE = (256, 32, 8)      # TrainingStepVerified(256, 32, 8, 1000): 976 k wires, domain 2^20 (BASELINE configs[4])
# SURVEY 8(d) normalisation: 136 MAC per 8-limb Montgomery product, 1360 MAC per G1 mixed add, 16 windows (c = 16)
MAC_PER_G1_POINT = 16 * 1360
MAC_PER_G2_POINT = 16 * 4080
N_SM, WIDE_MAC_PER_CLK_PER_SM = 148, 32      # nominal: IMAD.WIDE issues at a quarter of the 128-lane rate
PUBLISHED_PROOFS_PER_S = 0.147   # BASELINE.md section 1 (derived from Report.pdf Table 3, i7-10750H, snarkjs CLI)


def log(msg):
    print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zkfl", choices=["zkfl", "reference"])
    ap.add_argument("--workload", default="proofs", choices=["proofs", "split"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("ZKFL_BENCH_BATCH", "1024")))
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("ZKFL_BENCH_LANES", "2")),
                    help="contexts (streams) per GPU; the batch of a step is split evenly over them and proved concurrently")
    ap.add_argument("--no-msm", action="store_true", help="skip the standalone 2^20-point G1 MSM measurement")
    ap.add_argument("--no-split", action="store_true", help="skip the 2^20-domain split-proof measurement")
    ap.add_argument("--no-round", action="store_true", help="skip the 1 023-client full-round measurement (BASELINE configs[3])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs (dev runs under a profiler)")
    return ap.parse_args()


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth_inputs(prover, cc, batch: int, rank: int):
    """B DISTINCT synthetic clients per rank (the same count at every N): data from the reference's seeded generator
    (full_system_simulation.mjs client data), non-zero weights as in test_verified_gradient.mjs:230-235; the Poseidon commitments
    of all clients come from the GPU commitment pipeline in one batched pass (commitments.hydrate)."""
    import random
    from zkfl_b200 import commitments, inputs
    from zkfl_b200.formats import FR
    lcg = inputs.JsLcg(12345 + 1000003 * rank)
    clients, weights = [], []
    for i in range(batch):
        cl = inputs.SimClient(rank * batch + i + 1, lcg, hashed=False)
        cl.TAU2 = 1 << 62
        clients.append(cl)
        weights.append([lcg.random_int(-1000, 999) for _ in range(cl.DIM)])
    commitments.hydrate(prover, clients, weights)
    ins = b"".join(b"".join(int(v).to_bytes(32, "little") for v in cc.flatten_input(cl.training_input(w)))
                   for cl, w in zip(clients, weights))
    rnd = random.Random(1000 + rank)
    rs = b"".join(rnd.randrange(FR).to_bytes(32, "little") for _ in range(2 * batch))
    return ins, rs


def roofline_peaks(prover, sm_mhz):
    """live integer-pipe rates on this GPU (the denominators), and the nominal wide-MAC peak at the sampled SM clock"""
    imad = prover.bench_imad(148 * 2048 * 4, 4096)
    mac = prover.bench_widemac(148 * 2048 * 4, 4096)
    modmul = prover.bench_modmul(148 * 2048, 512)
    nominal = N_SM * WIDE_MAC_PER_CLK_PER_SM * (sm_mhz or 1965) * 1e6
    return {"widemac_per_s": mac, "imad32_per_s": imad, "modmul_per_s": modmul, "nominal_widemac_per_s": nominal}


def mac_roofline(kernel, mac, ms, peaks, extra=None):
    achieved = mac / (ms * 1e-3)
    out = {"bound": "imad", "kernel": kernel, "achieved": achieved / 1e9, "peak": peaks["widemac_per_s"] / 1e9, "unit": "GMAC/s",
           "frac": achieved / peaks["widemac_per_s"], "peak_nominal": peaks["nominal_widemac_per_s"] / 1e9,
           "frac_nominal": achieved / peaks["nominal_widemac_per_s"], "ms": ms}
    out.update(extra or {})
    return out


def bench_msm_2pow20(prover, torch, peaks, n: int = 1 << 20, reps: int = 5):
    """BASELINE.json's second metric: one G1 MSM over 2^20 RESIDENT points (bases k_i*G made by the device generator kernel;
    zkfl_msm_bases_load builds the window-shifted table once), uniform 254-bit scalars, fixed seed.
    value: scalars resident; e2e: scalars copied from pinned host memory each run."""
    import random
    from zkfl_b200.formats import FR
    rnd = random.Random(2020)
    bases = prover.g1_mul_generator(b"".join(rnd.randrange(FR).to_bytes(32, "little") for _ in range(n)))
    sc = b"".join(rnd.randrange(FR).to_bytes(32, "little") for _ in range(n))
    pin = torch.empty(len(sc), dtype=torch.uint8).pin_memory()
    pin.copy_(torch.frombuffer(bytearray(sc), dtype=torch.uint8))
    out = torch.empty(64, dtype=torch.uint8).pin_memory()
    h = prover.msm_load_bases(bases, 1)
    for _ in range(3):
        prover.msm_run(h, pin.data_ptr(), n, out.data_ptr())
    first = bytes(out.numpy().tobytes())
    prover.timer_begin()
    for _ in range(reps):
        prover.msm_run(h, None, n, None)
    ms = prover.timer_end() / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        prover.msm_run(h, pin.data_ptr(), n, out.data_ptr())
    e2e_ms = (time.perf_counter() - t0) * 1e3 / reps
    assert bytes(out.numpy().tobytes()) == first
    prover.prof_enable(True)
    prover.msm_run(h, None, n, out.data_ptr())
    stages = {k: round(v["ms"], 3) for k, v in prover.prof_read().items()}
    prover.prof_enable(False)
    prover.msm_free_bases(h)
    return {"metric": "g1_msm_points_per_s_2pow20", "value": n / (ms * 1e-3), "unit": "points/s", "ms": ms, "stages_ms": stages,
            "e2e": {"value": n / (e2e_ms * 1e-3), "ms": e2e_ms, "h2d_bytes": len(sc), "d2h_bytes": 64}, "points": n,
            "roofline": mac_roofline("whole MSM (sort + accumulate + fix-up + reduce), SURVEY 8d normalisation n x 21 760 MAC",
                                     n * MAC_PER_G1_POINT, ms, peaks,
                                     {"accumulate_only": mac_roofline("k_msm_accumulate_chunks<Fq>", n * MAC_PER_G1_POINT,
                                                                      stages.get("msm_acc_g1", ms), peaks)}),
            "parity": "tests/test_gpu_parity.py::test_g1_msm_2pow20_resident_table_against_oracle"}


def cpu_arm(cc, zkey_bytes, circuit_pack, sample: int, nthreads: int):
    """times the oracle on `sample` proofs (witness + prove), all threads; returns proofs/s"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib as ol
    ol.build()
    from zkfl_b200.formats import FR
    import random
    rnd = random.Random(5)
    ins = b"".join(circuit_pack[i % len(circuit_pack)] for i in range(sample))
    rs = b"".join(rnd.randrange(FR).to_bytes(32, "little") for _ in range(2 * sample))
    t0 = time.perf_counter()
    wt = ol.witness_batch(cc.program_bytes(), ins, cc.n_inputs, cc.n_wires, nthreads)
    t1 = time.perf_counter()
    ol.groth16_prove_batch(zkey_bytes, wt, rs, nthreads)
    t2 = time.perf_counter()
    return sample / (t2 - t0), (t1 - t0) * 1e3 / sample, (t2 - t1) * 1e3 / sample


def run_reference(args, rank, world):
    if rank != 0:
        return
    import zkfl_b200  # noqa: F401
    from zkfl_b200 import inputs
    from zkfl_b200.circuits import build_circuit
    cc = build_circuit("sgd_verified")
    # proving key made by the ORACLE alone (C++ scalar multiplications): nothing of the CUDA library is on this arm
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import groth16_ref
    import witness_ref
    from zkfl_b200.zkey_setup import toxic_from_seed
    zk = groth16_ref.setup_fast(witness_ref.R1cs(cc.r1cs_bytes()), *toxic_from_seed(b"zkfl-bench"))
    ncores = len(os.sched_getaffinity(0))
    objs = inputs.sgd_verified_batch(4, seed=12345, nonzero_weights=True)
    packs = [b"".join(int(v).to_bytes(32, "little") for v in cc.flatten_input(o)) for o in objs]
    sample = max(ncores, 4)
    for _ in range(min(args.warmup, 1)):
        cpu_arm(cc, zk, packs, max(ncores // 4, 1), ncores)
    t0 = time.perf_counter()
    vals = [cpu_arm(cc, zk, packs, sample, ncores) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32x8 (254-bit modular integers)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_proofs_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": "port",
                         "sample": f"{sample} proofs per step (witness + prove), one proof per thread, C++ oracle restating snarkjs "
                                   f"(snarkjs itself cannot run: no Node.js on this image); amortised per proof: witness {vals[-1][1]:.1f} ms, prove {vals[-1][2]:.0f} ms"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def bench_split(prover, torch, dist, rank, world, barrier, reps: int = 5, with_cpu: bool = True):
    """BASELINE configs[4] / north_star: ONE proof of the scaled training circuit (domain 2^20) with every MSM split by point range
    over the `world` GPUs and the partial sums all-gathered device-to-device (NCCL over NVLink).  Every rank makes the same key from
    the same seed on its own GPU (zkfl_groth16_setup: no 1.6 GB broadcast).  Wall-clock latencies of the blocking calls, barrier +
    synchronize on both sides, max over ranks, median of `reps`."""
    from zkfl_b200 import commitments, inputs, sharding
    from zkfl_b200.circuits.library import training_step_verified
    BATCH, DIM, DEPTH = SPLIT_SHAPE
    t0 = time.perf_counter()
    cc = training_step_verified(BATCH, DIM, DEPTH, 1000, f"sgd_scaled_{BATCH}_{DIM}_{DEPTH}")
    t_build = time.perf_counter() - t0
    lcg = inputs.JsLcg(2024)
    cl = inputs.SimClient(1, lcg, n=BATCH, dim=DIM, depth=DEPTH, hashed=False)
    cl.TAU2 = 1 << 62
    w = [lcg.random_int(-1000, 999) for _ in range(DIM)]
    commitments.hydrate(prover, [cl], w)
    inp = cl.training_input(w)
    t0 = time.perf_counter()
    zk = prover.new_zkey(cc, b"zkfl-bench-split")         # fixed seed: every rank derives the same (benchmark-only) key
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    Z = prover.load_zkey(zk, nparts=world)                 # window tables sized for a rank's share of the points
    circ = prover.load_circuit(cc, check_constraints=False)
    t_load = time.perf_counter() - t0
    packed = circ.pack_inputs([inp])
    rs = [(3, 4)]

    def timed(fn):
        out, ts = None, []
        for i in range(reps + 2):
            barrier()
            t = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if i >= 2:                      # the first calls allocate the workspace
                ts.append(float(dt.item()))
        return out, sorted(ts)[len(ts) // 2] * 1e3

    if dist is not None:
        full, t_full = timed(lambda: sharding.full_prove_split(prover, circ, Z, packed, rs, check=False))
        prover.witness_resident(circ, packed, check=False)
        only, t_prove = timed(lambda: sharding._exchange_and_finalize(prover, Z, None, 1, rs))
    else:
        full, t_full = timed(lambda: prover.full_prove(circ, Z, packed, rs, check=False)[0])
        prover.witness_resident(circ, packed, check=False)
        only, t_prove = timed(lambda: prover.finalize(Z, [prover.msm_partials(Z, None, 0, 1, B=1)], 1, rs))
    assert full == only, "split proof with a resident witness differs"
    prover.prof_enable(True)
    if dist is not None:
        sharding.full_prove_split(prover, circ, Z, packed, rs, check=False)
    else:
        prover.full_prove(circ, Z, packed, rs, check=False)
    stages = {k: round(v["ms"], 3) for k, v in prover.prof_read().items()}
    prover.prof_enable(False)
    res = None
    if rank == 0:
        whole = prover.full_prove(circ, Z, packed, rs, check=False)      # the whole proof on this one GPU: the split must equal it
        from zkfl_b200 import formats
        from zkfl_b200 import snarkjs as sj
        verified = sj.groth16.verify(formats.export_verification_key(zk), formats.publics_bytes_to_json(whole[1][0]),
                                     formats.proof_bytes_to_json(full[0]))
        cpu = None
        if with_cpu:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle_lib as ol
            ws = prover.calculate_witness(circ, packed, check=False)
            t = time.perf_counter()
            ref_p, _ = ol.groth16_prove(zk, ws[0], 3, 4)
            cpu = {"seconds": time.perf_counter() - t, "threads": ol.ncores(), "kind": "port", "bit_exact": ref_p == full[0],
                   "sample": "this one proof (prove step, witness given), C++ oracle on all host threads"}
        limit = max(((k, v) for k, v in stages.items() if not k.startswith("msm_reduce")), key=lambda kv: kv[1])
        res = {"circuit": f"TrainingStepVerified({BATCH},{DIM},{DEPTH},1000)", "n_wires": cc.n_wires, "n_constraints": cc.n_constraints,
               "domain": Z.domain, "n_gpus": world, "scaling": "strong",
               "prove_ms": t_prove, "full_prove_ms": t_full,
               "note": "prove_ms: witness resident in HBM -> proof bytes on the host (the `groth16 prove` step); full_prove_ms: circuit inputs "
                       "on the host -> witness on every GPU -> proof bytes on the host; every MSM split by point range, partial sums "
                       "all-gathered device-to-device, wall clock, max over ranks",
               "stages_ms_rank0": stages, "largest_main_stream_stage": {"name": limit[0], "ms": limit[1]},
               "equals_single_gpu_proof": whole[0][0] == full[0], "verified": bool(verified), "cpu_baseline": cpu,
               "one_off_s": {"circuit_build": round(t_build, 2), "setup_on_gpu": round(t_setup, 2), "key_load_and_tables": round(t_load, 2)},
               "zkey_bytes": len(zk)}
    Z.close()
    circ.close()
    return res


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import zkfl_b200  # noqa: F401
    from zkfl_b200.api import Prover
    from zkfl_b200.circuits import build_circuit

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: zkfl_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner to stdout when the first communicator is created; stdout must carry
        # exactly one JSON line, so point fd 1 at stderr while the process group comes up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    B, K, W = args.batch, args.steps, max(args.warmup, 0)
    cc = build_circuit("sgd_verified")
    lanes = max(1, min(args.lanes, B))
    provers = [Prover(local_rank) for _ in range(lanes)]
    prover = provers[0]
    # setup is per circuit, not per step: on the GPU, same (benchmark-only) seed on every rank
    zk = prover.new_zkey(cc, b"zkfl-bench")
    # the compiled program and the proving key (coefficients, window-shifted base tables: ~100 MB) are read-only device
    # data: ONE resident copy is shared by all contexts of this GPU, so the tables stay L2-resident
    circuit = prover.load_circuit(cc)      # with its R1CS: the `===` check runs on the device inside every proving pass, as fullProve does
    zkey = prover.load_zkey(zk)
    circuits, zkeys = [circuit] * lanes, [zkey] * lanes
    t_in = time.perf_counter()
    ins, rs = synth_inputs(prover, cc, B, rank)
    t_in = time.perf_counter() - t_in
    n_distinct = len({ins[i * 32 * circuit.n_inputs:(i + 1) * 32 * circuit.n_inputs] for i in range(B)})
    # lane k proves proofs [lo_k, hi_k) of the step's batch
    bounds = [(B * k // lanes, B * (k + 1) // lanes) for k in range(lanes)]
    in_sz, l = 32 * circuit.n_inputs, zkey.n_public

    # pinned host buffers for the end-to-end arm
    pin_in = torch.empty(len(ins), dtype=torch.uint8).pin_memory()
    pin_in.copy_(torch.frombuffer(bytearray(ins), dtype=torch.uint8))
    pin_rs = torch.empty(len(rs), dtype=torch.uint8).pin_memory()
    pin_rs.copy_(torch.frombuffer(bytearray(rs), dtype=torch.uint8))
    pin_proofs = torch.empty(256 * B, dtype=torch.uint8).pin_memory()
    pin_pubs = torch.empty(32 * l * B, dtype=torch.uint8).pin_memory()
    l2_flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def run_resident():
        for k, p in enumerate(provers):
            p.run_staged(circuits[k], zkeys[k], bounds[k][1] - bounds[k][0], check=True)

    def fetch_all():
        for k, p in enumerate(provers):
            lo, hi = bounds[k]
            p.fetch(hi - lo, pin_proofs.data_ptr() + 256 * lo)

    def e2e_step():
        def one(k):
            lo, hi = bounds[k]
            provers[k].full_prove_raw(circuits[k], zkeys[k], pin_in.data_ptr() + in_sz * lo, pin_rs.data_ptr() + 64 * lo, hi - lo,
                                      pin_proofs.data_ptr() + 256 * lo, pin_pubs.data_ptr() + 32 * l * lo)
        if lanes == 1:
            one(0)
            return
        th = [threading.Thread(target=one, args=(k,)) for k in range(lanes)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    log(f"setup done: B={B} lanes={lanes} world={world}, {n_distinct} distinct inputs generated in {t_in:.1f} s")
    # ---- resident arm (value)
    for k, p in enumerate(provers):
        lo, hi = bounds[k]
        p.stage(circuits[k], zkeys[k], pin_in.data_ptr() + in_sz * lo, pin_rs.data_ptr() + 64 * lo, hi - lo)
    for _ in range(W):
        run_resident()
    fetch_all()
    first_all = bytes(pin_proofs.numpy().tobytes())
    first = first_all[:256]
    barrier()
    launches0 = prover.launch_count()
    with ClockSampler(local_rank) as clocks:
        step_ms = []
        for _ in range(K):
            l2_flush.zero_()
            torch.cuda.synchronize()
            for p in provers[1:]:
                p.fetch(1, None)              # drain the other lanes (synchronises their streams)
            prover.timer_begin()
            run_resident()
            for p in provers[1:]:
                prover.wait_other(p)          # lane 0's stream joins the others before the end timestamp
            step_ms.append(prover.timer_end())
    launches = prover.launch_count() - launches0
    barrier()
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    fetch_all()
    assert bytes(pin_proofs.numpy().tobytes()) == first_all, "non-deterministic proofs for fixed r, s"

    log(f"resident arm done: {total_ms / K:.1f} ms/step")
    # ---- end-to-end arm (host buffers, H2D + D2H inside the timed region)
    for _ in range(min(W, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    assert bytes(pin_proofs.numpy().tobytes()) == first_all

    log(f"e2e arm done: {e2e_s * 1e3 / K:.1f} ms/step")
    prof = prof_step_ms = clk = peaks = msm = None
    if rank == 0:
        # ---- per-stage profile of one more step (CUDA events on the library's stream)
        # (a full batch on one context, so the stage times below are for B proofs without lane overlap)
        prover.stage(circuit, zkey, pin_in.data_ptr(), pin_rs.data_ptr(), B)
        prover.run_staged(circuit, zkey, B, check=True)
        prover.fetch(1, None)
        prover.prof_enable(True)
        prover.timer_begin()
        prover.run_staged(circuit, zkey, B, check=True)
        prof_step_ms = prover.timer_end()
        prof = prover.prof_read()
        prover.prof_enable(False)
        clk = clocks.summary()
        peaks = roofline_peaks(prover, clk.get("sm_mhz"))
        msm = bench_msm_2pow20(prover, torch, peaks) if not args.no_msm else None
    # ---- the batch verifier on this step's proofs (reported beside the headline, not part of it)
    verify = None
    if rank == 0:
        from zkfl_b200.formats import export_verification_key, vkey_json_to_bytes
        vkb = vkey_json_to_bytes(export_verification_key(zk))
        all_p, all_q = bytes(pin_proofs.numpy().tobytes()), bytes(pin_pubs.numpy().tobytes())
        vp = [all_p[256 * b:256 * (b + 1)] for b in range(B)]
        vq = [all_q[32 * l * b:32 * l * (b + 1)] for b in range(B)]
        prover.verify_batch(vkb, vq[:8], vp[:8])
        prover.verify_batch(vkb, vq, vp)                       # first full-size call allocates the workspace
        t0 = time.perf_counter()
        vok = prover.verify_batch(vkb, vq, vp)
        verify_ms = (time.perf_counter() - t0) * 1e3
        verify = {"proofs": B, "all_valid": bool(all(vok)), "ms": verify_ms, "proofs_per_s": B / (verify_ms * 1e-3),
                  "note": "zkfl_groth16_verify_batch on this step's proofs, host buffers, wall clock of the blocking call"}
    # ---- one large proof split over the GPUs of this run (every rank takes part)
    split = None
    if not args.no_split or args.workload == "split":
        for p in provers[1:]:
            p.close()
        split = bench_split(prover, torch, dist, rank, world, barrier, with_cpu=not args.no_cpu and world == 1)
        log("split proof done" + (f": prove {split['prove_ms']:.1f} ms, full prove {split['full_prove_ms']:.1f} ms" if split else ""))
    line = None
    if rank == 0:
        m, n, l = zkey.n_vars, zkey.domain, zkey.n_public
        from zkfl_b200.formats import read_container
        sec = read_container(zk, b"zkey")
        n_b2 = sum(1 for i in range(m) if sec[6][64 * i:64 * i + 64] != bytes(64))     # wires with a B-query point (the others are dropped)
        g1_pts = B * (2 * m - l - 1 + n_b2 + n)        # A (m), C (m - l - 1), B1 (n_b2), H (n)
        acc_ms, acc2_ms, ntt_ms = prof["msm_acc_g1"]["ms"], prof["msm_acc_g2"]["ms"], prof["ntt"]["ms"]
        lg = n.bit_length() - 1
        ntt_products = B * (6 * (n // 2) * lg + 3 * n + n)        # six transforms, the folded n^-1 / coset scaling, joinABC
        measured, traffic = {}, {}
        try:
            measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        except OSError:
            pass
        # reductions (side streams) and sorts (sort stream) overlap the product-bound stages of the main stream
        main_stream = sum(v["ms"] for k, v in prof.items() if not k.startswith(("msm_reduce", "msm_sort")))
        cpu = None
        if not args.no_cpu and world == 1:
            # ---- CPU baseline (bounded sample, all host cores; at N = 1 only: with more ranks the other processes hold the cores)
            ncores = len(os.sched_getaffinity(0))
            packs = [ins[i * 32 * circuit.n_inputs:(i + 1) * 32 * circuit.n_inputs] for i in range(min(B, 4))]
            sample = max(ncores, 4)
            cpu_val, cpu_w_ms, cpu_p_ms = cpu_arm(cc, zk, packs, sample, ncores)
            cpu = {"value": cpu_val, "unit": UNIT, "cores": ncores, "kind": "port",
                   "sample": f"{sample} proofs, one proof per thread (amortised per proof: witness {cpu_w_ms:.1f} ms + prove {cpu_p_ms:.0f} ms); "
                             "C++ oracle restating snarkjs (snarkjs cannot run: no Node.js on this image)"}
        # verify one proof of the batch with the oracle's pairing check (outside any timed region)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import groth16_ref
        import oracle_lib
        vk = groth16_ref.vkey_from_json(export_verification_key(zk))
        pubs0 = oracle_lib.ints(bytes(pin_pubs.numpy()[:32 * l].tobytes()))
        verified = groth16_ref.verify(vk, pubs0, groth16_ref.proof_from_bytes(first))
        value = world * B * K / (total_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": value / PUBLISHED_PROOFS_PER_S,
            "baseline_note": "BASELINE.md: 0.147 proofs/s = 1 / 6.8 s, snarkjs CLI `groth16 prove` of this circuit on an i7-10750H "
                             "laptop (Report.pdf Table 3) -- the reference's only published figure; different hardware, CPU only; not a headline",
            "dtype": "u32x8 (254-bit modular integers)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "lanes_per_gpu": lanes, "n_vars": m, "domain": n, "n_public": l,
                       "distinct_inputs_per_gpu": n_distinct, "constraint_check": "on the device inside every pass (circom's === semantics)",
                       "l2": "flushed (192 MB write) between timed steps",
                       "sharding": "independent proofs, b -> rank, no collective", "proof_verified_by_oracle": bool(verified)},
            "e2e": {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": len(ins) + len(rs),
                    "d2h_bytes_per_step": 256 * B + 32 * l * B},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": mac_roofline("k_msm_accumulate_chunks<Fq> (4 launches per step: A, C, B1, H)", g1_pts * MAC_PER_G1_POINT, acc_ms, peaks, {
                # per launch, like `achieved`: one A-query launch covers the B / lanes proofs of one context
                "traffic": (traffic["g1_accumulate_dram_bytes_per_proof"] * (B // lanes)) if traffic else None,
                "traffic_source": traffic.get("source"),
                "algorithmic_bytes_per_launch": (traffic["g1_accumulate_algorithmic_bytes_per_proof"] * (B // lanes)) if traffic else None,
                "proofs_per_launch": B // lanes,
                "share_of_step": acc_ms / prof_step_ms, "share_of_main_stream": acc_ms / main_stream,
                "peak_imad32_gops": peaks["imad32_per_s"] / 1e9, "modmul_per_s": peaks["modmul_per_s"],
                "note": "MAC = 32x32->64 multiply-accumulate; algorithmic MAC = G1 points x 16 windows x 1360 (SURVEY 8d normalisation); "
                        "peak = live microbenchmark of fused wide MACs (zkfl_bench_widemac, IMAD.WIDE.U32.X, profiles/r02_widemac_sass.txt); "
                        "peak_nominal = 148 SMs x 32 per clock x the SM clock sampled during the timed region; share_of_step: against the "
                        "device time of the profiled step (one context, B proofs)"}),
            "roofline_g2": mac_roofline("k_msm_accumulate_chunks<Fq2> (B2 query)", B * n_b2 * MAC_PER_G2_POINT, acc2_ms, peaks,
                                        {"share_of_step": acc2_ms / prof_step_ms}),
            "roofline_ntt": {"bound": "imad", "kernel": "H polynomial: 3 iNTT + coset + 3 NTT (radix-8 passes) + joinABC",
                             "achieved": ntt_products / (ntt_ms * 1e-3), "peak": peaks["modmul_per_s"], "unit": "Montgomery products/s",
                             "frac": ntt_products / (ntt_ms * 1e-3) / peaks["modmul_per_s"], "ms": ntt_ms,
                             "hbm_gbs": B * 512 * n / (ntt_ms * 1e-3) / 1e9, "hbm_peak_gbs": measured.get("hbm_gbs"),
                             "note": "at n = 2^14 the stage is bound by its Montgomery products (peak = live zkfl_bench_modmul rate), not by HBM: "
                                     "its algorithmic 512 n B per proof (SURVEY 8d) run at hbm_gbs of hbm_peak_gbs"},
            "stages_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
            "stages_note": "one context, B proofs; msm_reduce_* run on side streams and overlap the next accumulation, msm_sort_* run on the sort "
                           "stream beside A.w/B.w, the NTTs and the previous accumulations (their times are stream time, not additional step "
                           "time); profiled_step_ms is the device time of that step",
            "profiled_step_ms": prof_step_ms,
            "msm_g1_2pow20": msm,
            "split_proof": split,
            "verify_batch": verify,
            "full_round_1023": None,
            "cpu_baseline": cpu,
        }
        if args.workload == "split" and split is not None:
            line = {"metric": "groth16_full_prove_latency_ms_2pow20_domain", "value": split["full_prove_ms"], "unit": "ms", "n_gpus": world,
                    "steps": K, "warmup": W, "ms_per_step": split["full_prove_ms"], "higher_is_better": False, "scaling": "strong",
                    "vs_baseline": None, "dtype": "u32x8 (254-bit modular integers)", "data": "synthetic",
                    "config": {"workload": f"ONE proof of {split['circuit']} (domain 2^20) split over the GPUs of the run"},
                    "e2e": {"value": split["full_prove_ms"], "unit": "ms", "h2d_bytes_per_step": 32 * circuit.n_inputs, "d2h_bytes_per_step": 256 + 32 * 6},
                    "gpu_launches": int(launches), "clocks": clk, "split_proof": split, "cpu_baseline": split.get("cpu_baseline"),
                    "roofline": line["roofline"], "proofs_per_s_line": {"value": value, "unit": UNIT}}
    # ---- BASELINE configs[3]: the reference's whole round (tests/full_system_simulation.mjs:1244-1395) at 1 023 clients over the GPUs
    # of this run: inputs from the GPU commitment pipeline, 3 x 1 023 proofs sharded b -> rank, batch verification and the masked
    # aggregation on rank 0's GPU.  Second of two rounds (the first makes the three keys and allocates the workspaces).
    # A SIDE measurement, made after the headline line is complete: it must never cost that line.  If a rank fails inside one of
    # its collectives the others would wait for it forever, so every rank arms a watchdog that prints the line (rank 0) and leaves.
    round_failed = False
    if not args.no_round and args.workload == "proofs":
        import threading

        def _bail():
            try:
                if line is not None:
                    line["full_round_1023"] = {"error": "timeout: the round did not finish in 300 s"}
                    print(json.dumps(line), flush=True)
            finally:
                os._exit(0)
        wd = threading.Timer(300.0, _bail)
        wd.daemon = True
        wd.start()
        full_round = None
        try:
            from zkfl_b200 import simulation
            cache = {}
            for p in provers[1:]:
                p.close()
            simulation.run_round(prover, 1023, cache=cache, setup_seed=b"zkfl-bench-round")
            barrier()
            rep = simulation.run_round(prover, 1023, cache=cache, setup_seed=b"zkfl-bench-round")
            barrier()
            if rep is not None:
                full_round = {"clients": rep["clients"], "n_gpus": rep.get("n_gpus", world), "proofs": rep.get("proofs"),
                              "verified": rep["verified"], "timing_s": {k: round(v, 4) for k, v in rep["timing"].items()},
                              "note": "one round through the Python host API (simulation.run_round), wall clock on rank 0, inputs included"}
        except Exception as e:
            full_round = {"error": f"{type(e).__name__}: {e}"}
            round_failed = True
        wd.cancel()
        log("full round done" + (f": {full_round['timing_s'].get('round_s')} s" if full_round and "timing_s" in full_round else ""))
        if line is not None:
            line["full_round_1023"] = full_round
    if line is not None:
        print(json.dumps(line), flush=True)
        log("line printed")
    if round_failed:          # the other ranks may be stuck in a collective of the failed side measurement: no barrier, just leave
        sys.stdout.flush()
        os._exit(0)
    barrier()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
