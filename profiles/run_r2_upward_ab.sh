# A/B of the query's look-above scan: warp-cooperative (default) against per-lane with opaque registers (MSIM_QUERY_UPWARD=lanes)
mkdir -p gpurun_out
for v in coop lanes; do
  MSIM_QUERY_UPWARD=$v timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-e2e-variants > gpurun_out/r2v_$v.json 2> gpurun_out/r2v_$v.err
  python - <<PY
import json
p=json.load(open("gpurun_out/r2v_$v.json")); print("$v", round(p["ms_per_step"]*1e3,1), "us/tick", p["config"]["counts_check"]["status"], [(k["name"],k["launches"],round(k["avg_us"],1)) for k in p["kernels"][:7]])
PY
done
MSIM_QUERY_UPWARD=lanes timeout 200 python profiles/stress_bands_1gpu.py --ticks 4000 --events 2 > gpurun_out/stress_lanes.log 2>&1; grep -v "^   flags" gpurun_out/stress_lanes.log | tail -4 | cut -c1-500
