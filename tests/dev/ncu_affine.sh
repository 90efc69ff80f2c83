set -x; mkdir -p gpurun_out
export ZKFL_MSM_AFFINE=1 ZKFL_BENCH_BATCH=1024 ZKFL_BENCH_LANES=1
CMD="python bench.py --steps 1 --warmup 1 --no-msm"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate_affine -s 5 -c 4 -o gpurun_out/prof_affine $CMD > gpurun_out/ncu_affine.log 2>&1
tail -3 gpurun_out/plain.log; tail -5 gpurun_out/ncu_affine.log
