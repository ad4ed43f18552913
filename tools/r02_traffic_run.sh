set -x
# per-launch DRAM traffic of the k = 1 sweep at the shard shapes of 2 / 4 / 8 GPUs (one GPU holds markers_per_gpu markers of all 200k samples)
for M in 250000 125000 62500; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pk2_stream -s 4 -c 2 --csv --log-file gpurun_out/r02_traffic_M$M.csv python tools/sweep_bench.py 200000 $M 1 > gpurun_out/ncu_traffic_$M.log 2>&1
  tail -8 gpurun_out/r02_traffic_M$M.csv
done
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_6.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_6.log); tail -5 gpurun_out/r02_gputests_6.log
