set -x
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-flags-only"
for v in 1 4; do MSIM_SCAN_VARIANT=$v $B > gpurun_out/ab_scan$v.json 2>/dev/null; done
python profiles/compare_bench.py gpurun_out/ab_scan1.json gpurun_out/ab_scan4.json
MSIM_SCAN_VARIANT=4 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharding.py -m gpu -x -q 2>&1 | tail -2
