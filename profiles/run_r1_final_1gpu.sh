set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; python profiles/show_bench.py gpurun_out/final_bench_1gpu.json
python bench.py --workload munich_1m_nocollisions > gpurun_out/final_bench_1m.json 2> gpurun_out/final_bench_1m.err; python profiles/show_bench.py gpurun_out/final_bench_1m.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 12 --warmup 3 --preroll 40 --no-cpu-baseline --e2e-steps 1 > gpurun_out/final_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"scan_tile|cell_scatter|query_kernel|move_kernel|arrive" -s 160 -c 12 -o gpurun_out/final_prof -f python bench.py --steps 40 --warmup 3 --preroll 40 --no-cpu-baseline --e2e-steps 1 > gpurun_out/final_ncu_full.log 2>&1; tail -2 gpurun_out/final_ncu_full.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"move_kernel|shard_emit|shard_integrate|shard_append|query_kernel" -s 330 -c 10 -o gpurun_out/final_prof_band -f python profiles/band_microbench.py 2500000 2 20 > gpurun_out/final_ncu_band.log 2>&1; tail -2 gpurun_out/final_ncu_band.log
python __graft_entry__.py smoke 2>&1 | tail -1
