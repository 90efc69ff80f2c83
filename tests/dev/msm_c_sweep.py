"""dev: window size of the resident-table G1 MSM over 2^20 points (ZKFL_MSM_C_TABLE), device time per run and stage split"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import zkfl_b200  # noqa: F401
from zkfl_b200.api import Prover
import bench
P = Prover(0)
peaks = bench.roofline_peaks(P, 1965)
for c in [int(x) for x in (sys.argv[1:] or ["16", "17", "18", "19", "20"])]:
    os.environ["ZKFL_MSM_C_TABLE"] = str(c)
    r = bench.bench_msm_2pow20(P, torch, peaks, reps=5)
    print(f"c={c}: {r['ms']:.3f} ms ({r['value'] / 1e6:.0f} M points/s), e2e {r['e2e']['ms']:.3f} ms", r["stages_ms"], flush=True)
