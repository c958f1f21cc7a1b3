# pipelined rebuild (scan + scatter of tick t+1 beside the query of tick t) against the serial rebuild (MSIM_PIPELINE_BUILD=0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest_gpu.log; tail -3 gpurun_out/r2z_pytest_gpu.log
for v in 1 0 1 0; do
  MSIM_PIPELINE_BUILD=$v timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-e2e-variants > gpurun_out/r2z_pipe$v.json 2> gpurun_out/r2z_pipe$v.err
  python - <<PY
import json
p=json.load(open("gpurun_out/r2z_pipe$v.json")); print("pipeline=$v", round(p["ms_per_step"]*1e3,1), "us/tick", p["config"]["counts_check"]["status"], [(k["name"],k["launches"],round(k["avg_us"],1)) for k in p["kernels"][:7]])
PY
done
timeout 200 python profiles/stress_bands_1gpu.py --ticks 3000 --events 2 > gpurun_out/stress_bands7.log 2>&1; grep -v "^   flags" gpurun_out/stress_bands7.log | tail -3 | cut -c1-500
