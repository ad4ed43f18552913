set -x
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_3.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_3.log); tail -15 gpurun_out/r02_gputests_3.log
python tools/sweep_bench.py 200000 500000 16,31 > gpurun_out/r02_sweep_wide7.txt 2>&1
SGB_DIGITS=5 python tools/sweep_bench.py 200000 500000 16,31 > gpurun_out/r02_sweep_wide5.txt 2>&1
cat gpurun_out/r02_sweep_wide7.txt gpurun_out/r02_sweep_wide5.txt
python tools/step2_bench.py 200000 16384 > gpurun_out/r02_step2_bench.txt 2>&1; cat gpurun_out/r02_step2_bench.txt
