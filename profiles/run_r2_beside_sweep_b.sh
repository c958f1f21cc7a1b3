mkdir -p gpurun_out
for v in "2 2" "1 1" "2 0"; do
  set -- $v
  MSIM_MOVE_BESIDE_CTAS=$1 MSIM_ARRIVE_BESIDE_CTAS=$2 timeout 100 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-e2e-variants --no-flags-only > gpurun_out/r2z_sweep_$1_$2.json 2> gpurun_out/r2z_sweep_$1_$2.err
  python - <<PY
import json
p=json.load(open("gpurun_out/r2z_sweep_$1_$2.json")); print("move $1 arrive $2:", round(p["ms_per_step"]*1e3,1), "us/tick", p["config"]["counts_check"]["status"], [(k["name"],round(k["avg_us"],1)) for k in p["kernels"][:6]])
PY
done
