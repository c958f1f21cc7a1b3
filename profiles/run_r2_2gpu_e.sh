# 2 GPUs: the multi-process exchange tests, the count-consistency tests, the bench at N = 1 and N = 2 in the driver's configuration
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_count_consistency.py -m gpu -x -q > gpurun_out/r2y_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest_2gpu.log; tail -3 gpurun_out/r2y_pytest_2gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-e2e-variants > gpurun_out/r2y_bench_1gpu.json 2> gpurun_out/r2y_bench_1gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 1 > gpurun_out/r2y_bench_2gpu.json 2> gpurun_out/r2y_bench_2gpu.err
python - <<PY
import json
for f in ("gpurun_out/r2y_bench_1gpu.json", "gpurun_out/r2y_bench_2gpu.json"):
    try:
        p=json.load(open(f)); c=p["config"]
        ks=p.get("kernels") and [(k["name"],round(k["avg_us"],1)) for k in p["kernels"][:7]] or c.get("kernel_us_per_step_rank0")
        print(f, round(p["ms_per_step"]*1e3,1), "us/tick", c["counts_check"]["status"], c.get("pairs_last_tick"), c.get("flagged_last_tick"), ks)
    except Exception as ex:
        print(f, "no line", ex)
PY
