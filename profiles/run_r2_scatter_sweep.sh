# CTAs per SM of the scatter while it shares the SMs with the previous tick's query (pipelined rebuild)
mkdir -p gpurun_out
for v in 8 6 4 3 2; do
  MSIM_SCATTER_BESIDE_CTAS=$v timeout 200 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-e2e-variants --no-flags-only > gpurun_out/r2sc_$v.json 2> gpurun_out/r2sc_$v.err
  python - <<PY
import json
p=json.load(open("gpurun_out/r2sc_$v.json")); print("scatter ctas/SM=$v", round(p["ms_per_step"]*1e3,1), "us/tick", p["config"]["counts_check"]["status"], [(k["name"],round(k["avg_us"],1)) for k in p["kernels"][:6]])
PY
done
