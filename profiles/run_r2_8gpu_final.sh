# 8-GPU lines of round 2 with the final kernels: BASELINE config 3 (the driver's SCALE configuration) at N = 8 and 4, configs 4 and 5,
# the collisions-off entity ranges (weak scaling), and config 3 with pass B behind the exchange (MSIM_SHARD_ARRIVE_EARLY=1)
mkdir -p gpurun_out
run() { N=$1; tag=$2; shift; shift; env $ENVV timeout ${T:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N "$@" > gpurun_out/r2f_${tag}_${N}gpu.json 2> gpurun_out/r2f_${tag}_${N}gpu.err; echo "$tag N=$N rc=$?"
  python - <<PY
import json
try:
    p=json.load(open('gpurun_out/r2f_${tag}_${N}gpu.json')); c=p['config']
    print('   us/tick', round(p['ms_per_step']*1e3,1), 'value %.4g' % p['value'], (c.get('counts_check') or {}).get('status'), 'pairs', c.get('pairs_last_tick'), 'flagged', c.get('flagged_last_tick'))
    print('   kernels rank0', c.get('kernel_us_per_step_rank0'))
except Exception as ex:
    print('   no line:', ex)
PY
}
run 8 munich10m --steps 20 --warmup 5 --e2e-steps 1
run 8 munich10m_b --steps 20 --warmup 5 --e2e-steps 1
ENVV="MSIM_SHARD_ARRIVE_EARLY=1" run 8 munich10m_early --steps 20 --warmup 5 --e2e-steps 1
ENVV="A=1" run 4 munich10m --steps 20 --warmup 5 --e2e-steps 1
ENVV="A=1" T=420 run 8 grid100m --workload grid4096_100m_collisions --steps 20 --warmup 5 --e2e-steps 1
ENVV="A=1" T=420 run 8 dense50m --workload munich_50m_dense --steps 20 --warmup 5 --e2e-steps 1
ENVV="A=1" run 8 nocoll_weak --workload munich_1m_nocollisions --entities 10000000 --scaling weak --steps 50 --warmup 5 --e2e-steps 1
