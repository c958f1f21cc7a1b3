#!/bin/bash
# Second GPU call of the next round (1 GPU, ~8 min): ncu evidence for whichever knobs the first call showed to pay.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash profiles/run_r2_ncu.sh "MSIM_ARRIVE_BESIDE_CTAS=2 MSIM_SCAN_MIN_BLOCKS=8" "--fused-arrive"'
# $1 = environment knobs, $2 = extra bench flags.  Launch list of two ticks (per-launch times, cold caches, serialised) and one
# --set full capture each of the move, scan, scatter, query and pass-B kernels; read here with
#   ncu -i gpurun_out/r2b_full.ncu-rep --page raw --csv | python profiles/summarize_ncu.py
set -x
mkdir -p gpurun_out
KNOBS="$1"; FLAGS="$2"
B="python bench.py --steps 40 --warmup 3 --preroll 40 --no-cpu-baseline --e2e-steps 1 $FLAGS"
env $KNOBS timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/r2b_launches.csv $B > gpurun_out/r2b_launches.log 2>&1
env $KNOBS timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:move_kernel|scan_tiles|cell_scatter|query|arrive_kernel' -s 60 -c 12 -o gpurun_out/r2b_full -f $B > gpurun_out/r2b_full.log 2>&1
tail -3 gpurun_out/r2b_full.log
python profiles/summarize_launches.py gpurun_out/r2b_launches.csv 2>/dev/null | head -30
