set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharding.py -m gpu -x -q > gpurun_out/r2k_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_2gpu.log
tail -5 gpurun_out/r2k_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 5 > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2k_bench_2gpu.json
