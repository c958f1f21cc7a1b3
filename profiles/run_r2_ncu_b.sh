# ncu evidence after the query / fold rewrite: launch list of two timed ticks, full captures of the two kernels that changed
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-flags-only --no-e2e-variants"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 140 --csv --log-file gpurun_out/r2b_launches.csv $B > gpurun_out/r2b_launches.log 2>&1
for k in query_tiles_kernel fold_counts_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 40 -c 1 -f -o gpurun_out/r2b_full_$k $B > gpurun_out/r2b_full_$k.log 2>&1
done
ls -la gpurun_out/r2b_full_* gpurun_out/r2b_launches.csv
