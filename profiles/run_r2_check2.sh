set -x
mkdir -p gpurun_out
python profiles/pcie_duplex.py > gpurun_out/r2l_pcie.json 2>&1; cat gpurun_out/r2l_pcie.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python profiles/compare_bench.py gpurun_out/r2l_bench.json
python -c "
import json; p=json.load(open('gpurun_out/r2l_bench.json')); c=p['config']
print({k:c[k] for k in ('move_passes_done','pairs_last_tick','flagged_last_tick','counts_check','resort')})
print(json.dumps(p['e2e'],indent=0))
"
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2l_ref.json 2> gpurun_out/r2l_ref.err ) 2>&1 | tail -4; echo "ref rc=$?"
cat gpurun_out/r2l_ref.json | cut -c1-1800
nproc
