mkdir -p gpurun_out
N=${N:-8}
run() { tag=$1; shift; env $ENVV timeout ${T:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N "$@" > gpurun_out/r2t_${tag}_${N}gpu.json 2> gpurun_out/r2t_${tag}_${N}gpu.err; echo "$tag rc=$?"
  python - <<PY
import json
try:
    p=json.load(open('gpurun_out/r2t_${tag}_${N}gpu.json')); c=p['config']
    print('$tag', 'us/tick', round(p['ms_per_step']*1e3,1), c.get('counts_check',{}).get('status'), 'pairs', c.get('pairs_last_tick'), 'flagged', c.get('flagged_last_tick'), c.get('pairs_flagged_by_rank'), c.get('splits'))
except Exception as ex:
    print('$tag no line:', ex)
PY
}
run a --steps 20 --warmup 5 --e2e-steps 1
run b --steps 20 --warmup 5 --e2e-steps 1
ENVV="MSIM_OVERLAP_TICKS=0" run serial --steps 20 --warmup 5 --e2e-steps 1
ENVV="A=1" run s10 --steps 10 --warmup 5 --e2e-steps 1
ENVV="A=1" run coll --steps 20 --warmup 5 --e2e-steps 1 --exchange collective
