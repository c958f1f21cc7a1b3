set -x
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-flags-only"
for v in 0 65536 262144 1048576; do MSIM_QUERY_PREFETCH=$v $B > gpurun_out/ab_pf$v.json 2>/dev/null; done
python profiles/compare_bench.py gpurun_out/ab_pf0.json gpurun_out/ab_pf65536.json gpurun_out/ab_pf262144.json gpurun_out/ab_pf1048576.json
