# dev: one full capture of the operand-file G2 accumulation at the benched size; CSV exports only (gpurun_out/ is capped at 64 MiB)
set -x; mkdir -p gpurun_out
export ZKFL_BENCH_LANES=1
CMD="python bench.py --steps 1 --warmup 1 --no-msm --no-split --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate_chunks_g2f -s 1 -c 1 -o /tmp/prof_g2f $CMD > gpurun_out/r2_ncu_g2f.log 2>&1
ncu -i /tmp/prof_g2f.ncu-rep --page raw --csv > gpurun_out/r02_g2f_raw.csv 2>/dev/null
ncu -i /tmp/prof_g2f.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_g2f_src.csv 2>/dev/null
gzip -f gpurun_out/r02_g2f_src.csv
tail -3 gpurun_out/r2_ncu_g2f.log
