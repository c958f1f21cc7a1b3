# 8-GPU lines of round 2: BASELINE config 3 (the driver's SCALE configuration), configs 4 and 5, and the collisions-off entity ranges
mkdir -p gpurun_out
N=${N:-8}
run() { tag=$1; shift; timeout ${T:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N "$@" > gpurun_out/r2s_${tag}_${N}gpu.json 2> gpurun_out/r2s_${tag}_${N}gpu.err; echo "$tag rc=$?"
  python - <<PY
import json
try:
    p=json.load(open('gpurun_out/r2s_${tag}_${N}gpu.json')); c=p['config']
    print('$tag', 'us/tick', round(p['ms_per_step']*1e3,1), 'value %.3g' % p['value'], c.get('counts_check'), 'pairs', c.get('pairs_last_tick'), 'owned', c.get('owned_per_rank'))
    print('   kernels rank0', c.get('kernel_us_per_step_rank0'))
except Exception as ex:
    print('$tag no line:', ex)
PY
}
run munich10m --steps 20 --warmup 5 --e2e-steps 1
T=420 run grid100m --workload grid4096_100m_collisions --steps 20 --warmup 5 --e2e-steps 1
T=420 run dense50m --workload munich_50m_dense --steps 20 --warmup 5 --e2e-steps 1
run nocoll_weak --workload munich_1m_nocollisions --entities 10000000 --scaling weak --steps 50 --warmup 5 --e2e-steps 1
nproc; free -g | head -2
