set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest_gpu.log
tail -5 gpurun_out/r2p_pytest_gpu.log
B="python bench.py --steps 64 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-flags-only --no-e2e-variants"
$B > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
MSIM_OVERLAP_TICKS=0 $B > gpurun_out/r2p_bench_nooverlap.json 2>/dev/null
MSIM_MOVE_BESIDE_CTAS=2 $B > gpurun_out/r2p_bench_b2.json 2>/dev/null
MSIM_MOVE_BESIDE_CTAS=3 $B > gpurun_out/r2p_bench_b3.json 2>/dev/null
MSIM_MOVE_BESIDE_CTAS=6 $B > gpurun_out/r2p_bench_b6.json 2>/dev/null
MSIM_MOVE_BESIDE_CTAS=8 $B > gpurun_out/r2p_bench_b8.json 2>/dev/null
MSIM_MOVE_BESIDE_CTAS=4 MSIM_ARRIVE_BESIDE_CTAS=2 $B > gpurun_out/r2p_bench_b4a2.json 2>/dev/null
MSIM_MOVE_BESIDE_CTAS=4 MSIM_ARRIVE_BESIDE_CTAS=4 $B > gpurun_out/r2p_bench_b4a4.json 2>/dev/null
python profiles/compare_bench.py gpurun_out/r2p_bench_nooverlap.json gpurun_out/r2p_bench.json gpurun_out/r2p_bench_b2.json gpurun_out/r2p_bench_b3.json gpurun_out/r2p_bench_b6.json gpurun_out/r2p_bench_b8.json gpurun_out/r2p_bench_b4a2.json gpurun_out/r2p_bench_b4a4.json
python -c "
import json
for f in ['r2p_bench','r2p_bench_nooverlap','r2p_bench_b8']:
    p=json.load(open('gpurun_out/%s.json'%f)); print(f, p['ms_per_step'], p['config']['counts_check']['status'])
"
