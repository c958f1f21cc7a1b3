#!/bin/bash
# First GPU call of the next round (1 GPU, ~25-30 min of box time; drop the gated pytest line to halve it):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash profiles/run_r2_first.sh'
# 1. the required suite (must stay green), 2. the gated tests of the paths written without a GPU, 3. one bench line per
# experiment knob (short runs: 100 steps, no CPU baseline), 4. collisions-off at 10 M with and without the fused pass B.
# Everything lands in gpurun_out/r2a_*.  Nothing here is a result until it has been read and copied to profiles/.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-flags-only"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2a_pytest_gpu.log
MSIM_TEST_UNVERIFIED=1 timeout 1500 python -m pytest tests/test_zz_gpu_unverified.py -m gpu -q > gpurun_out/r2a_pytest_unverified.log 2>&1; tail -15 gpurun_out/r2a_pytest_unverified.log

python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err  # with the colours-only extra
$B --fused-arrive > gpurun_out/r2a_bench_fused.json 2> gpurun_out/r2a_bench_fused.err
MSIM_MOVE_MIN_BLOCKS=5 $B > gpurun_out/r2a_bench_minb5.json 2> gpurun_out/r2a_bench_minb5.err
MSIM_MOVE_MIN_BLOCKS=6 $B > gpurun_out/r2a_bench_minb6.json 2> gpurun_out/r2a_bench_minb6.err
MSIM_MOVE_GRID=occupancy $B > gpurun_out/r2a_bench_occgrid.json 2> gpurun_out/r2a_bench_occgrid.err
for k in 1 2 4; do MSIM_ARRIVE_BESIDE_CTAS=$k $B > gpurun_out/r2a_bench_besidectas$k.json 2> gpurun_out/r2a_bench_besidectas$k.err; done
MSIM_L2_PERSIST_ROADS=1 $B > gpurun_out/r2a_bench_l2roads.json 2> gpurun_out/r2a_bench_l2roads.err
MSIM_QUERY_PAIRED=1 $B > gpurun_out/r2a_bench_paired.json 2> gpurun_out/r2a_bench_paired.err
MSIM_SCAN_MIN_BLOCKS=8 $B > gpurun_out/r2a_bench_scan8.json 2> gpurun_out/r2a_bench_scan8.err
MSIM_MOVE_MIN_BLOCKS=6 MSIM_MOVE_GRID=occupancy MSIM_SCAN_MIN_BLOCKS=8 MSIM_QUERY_PAIRED=1 $B --fused-arrive > gpurun_out/r2a_bench_all.json 2> gpurun_out/r2a_bench_all.err
MSIM_MOVE_MIN_BLOCKS=6 MSIM_MOVE_GRID=occupancy $B --fused-arrive > gpurun_out/r2a_bench_fused_minb6_occ.json 2> gpurun_out/r2a_bench_fused_minb6_occ.err
# collisions off: BASELINE configs[1] (1 M, L2 flushed between steps) and the same at 10 M (HBM-bound)
$B --workload munich_1m_nocollisions > gpurun_out/r2a_bench_1m_off.json 2> gpurun_out/r2a_bench_1m_off.err
$B --workload munich_1m_nocollisions --fused-arrive > gpurun_out/r2a_bench_1m_off_fused.json 2> gpurun_out/r2a_bench_1m_off_fused.err
$B --workload munich_1m_nocollisions --entities 10000000 > gpurun_out/r2a_bench_10m_off.json 2> gpurun_out/r2a_bench_10m_off.err
MSIM_L2_PERSIST_ROADS=1 $B --workload munich_1m_nocollisions --entities 10000000 > gpurun_out/r2a_bench_10m_off_l2roads.json 2> gpurun_out/r2a_bench_10m_off_l2roads.err
MSIM_ARRIVE_GRID=persistent $B --workload munich_1m_nocollisions --entities 10000000 > gpurun_out/r2a_bench_10m_off_persistent.json 2> gpurun_out/r2a_bench_10m_off_persistent.err
$B --workload munich_1m_nocollisions --entities 10000000 --fused-arrive > gpurun_out/r2a_bench_10m_off_fused.json 2> gpurun_out/r2a_bench_10m_off_fused.err
python profiles/compare_bench.py gpurun_out/r2a_bench_default.json gpurun_out/r2a_bench_[!d]*.json
$B --e2e-pipelined --e2e-steps 5 > gpurun_out/r2a_bench_e2e_pipelined.json 2> gpurun_out/r2a_bench_e2e_pipelined.err; python -c "import json; print(json.load(open(\"gpurun_out/r2a_bench_e2e_pipelined.json\"))[\"e2e\"])"
