# round-2 profiles: launch list of the default bench workload, then full captures of the bucket accumulation (G1 + G2) and of the
# fix-up / reduction kernels at the benched batch size.  Each ncu pass runs only after the same command exited 0 without ncu.
# The .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): raw / source pages are exported to CSV there.
set -x; mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-msm --no-split --no-cpu --no-round"
$CMD > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
export ZKFL_BENCH_LANES=1
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate_chunks -s 5 -c 5 -o /tmp/r02_prof_acc $CMD > gpurun_out/r2_ncu_acc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_reduce_level|k_msm_fixup|k_reduce_final" -s 20 -c 14 -o /tmp/r02_prof_red $CMD > gpurun_out/r2_ncu_red.log 2>&1
for r in acc red; do
  ncu -i /tmp/r02_prof_$r.ncu-rep --page raw --csv > gpurun_out/r02_${r}_raw.csv 2>/dev/null
done
# SASS-level source pages of the accumulation launches (A, C, B1 on G1; B2 on G2; H on G1)
for i in 0 3; do ncu -i /tmp/r02_prof_acc.ncu-rep --page source --csv --print-source sass --launch-skip $i --launch-count 1 > gpurun_out/r02_acc_src_$i.csv 2>/dev/null; done
gzip -f gpurun_out/r02_acc_src_*.csv
tail -2 gpurun_out/r2_plain.err; tail -2 gpurun_out/r2_ncu_acc.log; tail -2 gpurun_out/r2_ncu_red.log
du -sm gpurun_out
