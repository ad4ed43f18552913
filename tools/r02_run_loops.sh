set -x
(timeout 900 python -m pytest tests/test_gpu_driver_loops.py -m gpu -x -q > gpurun_out/r02_gputests_loops.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_loops.log); tail -25 gpurun_out/r02_gputests_loops.log
