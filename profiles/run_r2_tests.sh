set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest_gpu.log
tail -8 gpurun_out/r2q_pytest_gpu.log
