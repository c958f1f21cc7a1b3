# race hunt: totals over every collision pass, 1 GPU against N (profiles/stress_shard.py)
mkdir -p gpurun_out
N=${N:-4}; T=${T:-10000}
timeout 200 python profiles/stress_shard.py --ticks $T --every 100 > gpurun_out/stress_n1.log 2>&1; tail -2 gpurun_out/stress_n1.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR profiles/stress_shard.py --ticks $T --every 100 --modes p2p,p2p,p2p,collective > gpurun_out/stress_n$N.log 2>&1; grep -v "^W\|^\*\|OMP" gpurun_out/stress_n$N.log | tail -8
MSIM_OVERLAP_TICKS=0 timeout 300 $TR profiles/stress_shard.py --ticks $T --every 100 --modes p2p,p2p --tag _serial > gpurun_out/stress_n${N}_serial.log 2>&1; grep -v "^W\|^\*\|OMP" gpurun_out/stress_n${N}_serial.log | tail -5
