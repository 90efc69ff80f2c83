# one GPU call: full GPU test-suite, the default bench, launch list + full ncu capture of the dominant kernels
set -x; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json
CMD="python bench.py --batch 256 --lanes 1 --steps 1 --warmup 1 --no-msm"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r01_v4.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate_chunks -s 5 -c 5 -o gpurun_out/prof_chunks_v4 $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
