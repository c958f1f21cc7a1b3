# ncu evidence for round 2: launch list of two timed ticks, then one full capture per hot kernel (tick 300 of the bench's population)
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-flags-only --no-e2e-variants"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 120 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches.log 2>&1
for k in query_tiles_kernel cell_scatter_slots_kernel scan_cells_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 40 -c 1 -f -o gpurun_out/r2_full_$k $B > gpurun_out/r2_full_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:move_kernel --launch-skip 300 -c 1 -f -o gpurun_out/r2_full_move_kernel $B > gpurun_out/r2_full_move_kernel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:arrive_kernel --launch-skip 300 -c 1 -f -o gpurun_out/r2_full_arrive_kernel $B > gpurun_out/r2_full_arrive_kernel.log 2>&1
ls -la gpurun_out/r2_full_* gpurun_out/r2_launches.csv
python profiles/e2e_probe.py > gpurun_out/r2_e2e_probe.json 2> gpurun_out/r2_e2e_probe.err; cat gpurun_out/r2_e2e_probe.json
