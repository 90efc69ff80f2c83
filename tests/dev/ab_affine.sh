set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch_affine or large_batch or g1_msm_sizes or prove_sgd" 2>&1 | tail -5
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-msm > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["stages_ms"].items()})
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run aff_b1024_l1 ZKFL_MSM_AFFINE=1 ZKFL_BENCH_BATCH=1024 ZKFL_BENCH_LANES=1
run xyzz_b1024_l1 ZKFL_MSM_AFFINE=0 ZKFL_BENCH_BATCH=1024 ZKFL_BENCH_LANES=1
run aff_b2048_l1 ZKFL_MSM_AFFINE=1 ZKFL_BENCH_BATCH=2048 ZKFL_BENCH_LANES=1
run xyzz_b2048_l2 ZKFL_MSM_AFFINE=0 ZKFL_BENCH_BATCH=2048 ZKFL_BENCH_LANES=2
run aff_k32 ZKFL_MSM_AFFINE=1 ZKFL_MSM_AFFINE_K=32 ZKFL_BENCH_BATCH=2048 ZKFL_BENCH_LANES=1
if [ -n "$DO_NCU" ]; then
export ZKFL_MSM_AFFINE=1 ZKFL_BENCH_BATCH=1024 ZKFL_BENCH_LANES=1
CMD="python bench.py --steps 1 --warmup 1 --no-msm"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate_affine -s 5 -c 1 -o gpurun_out/prof_affine2 $CMD > gpurun_out/ncu_affine.log 2>&1
tail -2 gpurun_out/ncu_affine.log
fi
