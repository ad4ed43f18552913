set -x
(timeout 600 ncu --set full --import-source on --clock-control none -k regex:recomb_post1_wide -s 3 -c 1 -o gpurun_out/r02_post1_wide_k31 -f python tools/sweep_bench.py 200000 500000 31 > gpurun_out/r02_ncu_post1.log 2>&1); tail -3 gpurun_out/r02_ncu_post1.log
