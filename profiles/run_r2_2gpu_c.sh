set -x
mkdir -p gpurun_out
MSIM_SHARD_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 5 --e2e-steps 1 > gpurun_out/r2n_bench_2gpu.json 2> gpurun_out/r2n_bench_2gpu.err; echo "bench rc=$?"
grep "shard trace" gpurun_out/r2n_bench_2gpu.err
python -c "
import json; p=json.load(open('gpurun_out/r2n_bench_2gpu.json')); c=p['config']
print(p['ms_per_step'], c['kernel_us_per_step_by_rank'], c['counts_check']['status'], c['pairs_last_tick'], c['move_passes_done'])
"
