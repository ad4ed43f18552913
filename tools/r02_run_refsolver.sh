set -x
ls oracle/_ref
(timeout 900 python -m pytest tests/test_gpu_vs_reference_solver.py tests/test_gpu_driver_loops.py -m gpu -q > gpurun_out/r02_gputests_refsolver.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_refsolver.log); tail -15 gpurun_out/r02_gputests_refsolver.log
