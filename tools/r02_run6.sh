set -x
python tools/step2_bench.py 200000 16384 > gpurun_out/r02_step2_bench.txt 2>&1; cat gpurun_out/r02_step2_bench.txt
python tools/step2_bench.py 200000 16384 0.05 0.5 1e9 > gpurun_out/r02_step2_bench_nospa.txt 2>&1; cat gpurun_out/r02_step2_bench_nospa.txt
python tools/step2_bench.py 200000 16384 0.001 0.01 > gpurun_out/r02_step2_bench_rare.txt 2>&1; cat gpurun_out/r02_step2_bench_rare.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_step2_launches.csv python tools/step2_bench.py 200000 8192 > gpurun_out/ncu_s2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step2_kernel -s 3 -c 1 -o gpurun_out/r02_step2_kernel -f python tools/step2_bench.py 200000 8192 > gpurun_out/ncu_s2b.log 2>&1
ls -la gpurun_out/*.ncu-rep
