set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_gpu.log
tail -5 gpurun_out/r2j_pytest_gpu.log
python bench.py --steps 32 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python profiles/compare_bench.py gpurun_out/r2j_bench.json
