set -x
(timeout 1500 python -m pytest tests/ -m gpu -q > gpurun_out/r02_gputests_7.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_7.log); tail -8 gpurun_out/r02_gputests_7.log
