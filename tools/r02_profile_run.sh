set -x
free -g | head -2; df -h /dev/shm /tmp | tail -2; nproc
python tools/sweep_bench.py 200000 500000 1,2,3,4,8,16,31 > gpurun_out/r02_sweep_digits7.txt 2>&1
SGB_DIGITS=5 python tools/sweep_bench.py 200000 500000 3,4,8,16,31 > gpurun_out/r02_sweep_digits5.txt 2>&1
cat gpurun_out/r02_sweep_digits7.txt gpurun_out/r02_sweep_digits5.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/r02_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pk2_stream -s 4 -c 2 -o gpurun_out/r02_pk2_stream -f python tools/sweep_bench.py 200000 500000 1 > gpurun_out/ncu1.log 2>&1
SGB_DIGITS=5 ncu --set full --clock-control none --import-source on -k regex:pk2_umma -s 2 -c 2 -o gpurun_out/r02_pk2_umma_k31_d5 -f python tools/sweep_bench.py 200000 500000 31 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pk2_umma -s 2 -c 2 -o gpurun_out/r02_pk2_umma_k4 -f python tools/sweep_bench.py 200000 500000 4 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
