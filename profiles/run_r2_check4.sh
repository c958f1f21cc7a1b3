# after the query / fold rewrite: the whole GPU suite, the long stress, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest_gpu.log; tail -4 gpurun_out/r2w_pytest_gpu.log
timeout 200 python profiles/stress_bands_1gpu.py --ticks 5000 --events 2 > gpurun_out/stress_bands5.log 2>&1; grep -v "^   flags" gpurun_out/stress_bands5.log | tail -3 | cut -c1-500
timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-e2e-variants > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
python - <<PY
import json
p=json.load(open("gpurun_out/r2w_bench.json")); print(round(p["ms_per_step"]*1e3,1), "us/tick", p["config"]["counts_check"]["status"], [(k["name"],k["launches"],round(k["avg_us"],1)) for k in p["kernels"][:7]])
PY
