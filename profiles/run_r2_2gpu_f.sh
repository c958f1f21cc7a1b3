# 2 GPUs: pass B behind the exchange (default) against behind the scatter (MSIM_SHARD_ARRIVE_EARLY=0); sharding + consistency tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_count_consistency.py tests/test_gpu_configs45.py -m gpu -x -q > gpurun_out/r2y2_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y2_pytest_2gpu.log; tail -3 gpurun_out/r2y2_pytest_2gpu.log
for v in 1 0 1 0; do
MSIM_SHARD_ARRIVE_EARLY=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 64 --warmup 5 --e2e-steps 1 > gpurun_out/r2y2_bench_2gpu_$v.json 2> gpurun_out/r2y2_bench_2gpu_$v.err
python - <<PY
import json
f="gpurun_out/r2y2_bench_2gpu_$v.json"
try:
    p=json.load(open(f)); c=p["config"]
    print("early=$v", round(p["ms_per_step"]*1e3,1), "us/tick", c["counts_check"]["status"], c.get("pairs_last_tick"), c.get("flagged_last_tick"), c.get("kernel_us_per_step_rank0"))
except Exception as ex:
    print(f, "no line", ex)
PY
done
