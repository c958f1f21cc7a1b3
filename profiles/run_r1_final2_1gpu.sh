set -x
python bench.py > gpurun_out/final2_bench_1gpu.json 2> gpurun_out/final2_bench_1gpu.err; python profiles/show_bench.py gpurun_out/final2_bench_1gpu.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"query_kernel" -s 40 -c 2 -o gpurun_out/final2_prof_query -f python bench.py --steps 40 --warmup 3 --preroll 40 --no-cpu-baseline --e2e-steps 1 > gpurun_out/final2_ncu_query.log 2>&1; tail -2 gpurun_out/final2_ncu_query.log
