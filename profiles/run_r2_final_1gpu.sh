# last GPU call of round 2: the GPU suite, smoke(), and the bench line in the driver's configuration with everything on
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_final_pytest_gpu.log; tail -3 gpurun_out/r2_final_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
p=json.load(open("gpurun_out/r2_final_bench_1gpu.json")); c=p["config"]
print(round(p["ms_per_step"]*1e3,1), "us/tick", "%.4g" % p["value"], c["counts_check"]["status"], "e2e", p["e2e"]["value"], p["e2e"].get("variant"), "cpu", p["cpu_baseline"] and p["cpu_baseline"]["value"], p["roofline"]["frac"], p["tick"], p["clocks"])
print([(k["name"],k["launches"],round(k["avg_us"],1)) for k in p["kernels"][:8]])
print(p.get("move_only"), p.get("flags_only"))
PY
