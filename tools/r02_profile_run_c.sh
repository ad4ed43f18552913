set -x
ncu --set full --clock-control none --import-source on -k regex:dense_symm_mma -s 2 -c 1 -o gpurun_out/r02_dense_symm_mma_pairs_k4 -f python tools/dense_bench.py 40000 50000 7 full > gpurun_out/ncu_dm2.log 2>&1; tail -2 gpurun_out/ncu_dm2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense --no-ingest --no-step2 > gpurun_out/r02_bench_under_ncu2.log 2>&1; tail -2 gpurun_out/r02_bench_under_ncu2.log | cut -c1-300
ls -la gpurun_out/r02_dense_symm_mma_pairs_k4.ncu-rep gpurun_out/r02_launches_bench_final.csv
