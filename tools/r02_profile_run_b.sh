set -x
python tools/copy_bw.py > gpurun_out/r02_copy_bw.txt 2>&1; cat gpurun_out/r02_copy_bw.txt
ncu --set full --clock-control none --import-source on -k regex:dense_symm_mma -s 2 -c 1 -o gpurun_out/r02_dense_symm_mma_k4 -f python tools/dense_bench.py 40000 50000 7 full > gpurun_out/ncu_dm.log 2>&1; tail -3 gpurun_out/ncu_dm.log
SGB_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_step1_M62500.csv python tools/profile_step1_host.py 200000 62500 > gpurun_out/r02_step1_under_ncu.log 2>&1; tail -5 gpurun_out/r02_step1_under_ncu.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_step1_M62500.csv
