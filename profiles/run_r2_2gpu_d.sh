mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 5 --e2e-steps 1 > gpurun_out/r2r_$tag.json 2> gpurun_out/r2r_$tag.err
  python -c "
import json,sys; p=json.load(open('gpurun_out/r2r_$tag.json')); c=p['config']
print('$tag', round(p['ms_per_step']*1e3,1), c['kernel_us_per_step_rank0'], c['counts_check']['status'])"; }
run default A=1
run a2m4 MSIM_SHARD_ARRIVE_BESIDE_CTAS=2 MSIM_SHARD_MOVE_BESIDE_CTAS=4
run a4m8 MSIM_SHARD_ARRIVE_BESIDE_CTAS=4 MSIM_SHARD_MOVE_BESIDE_CTAS=8
run a1m2 MSIM_SHARD_ARRIVE_BESIDE_CTAS=1 MSIM_SHARD_MOVE_BESIDE_CTAS=2
run serial MSIM_OVERLAP_TICKS=0
