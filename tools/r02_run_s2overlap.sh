set -x
(timeout 900 python -m pytest tests/test_step2_batched.py tests/test_step2_golden.py tests/test_step2_formats.py tests/test_step2_rare_exact.py -m gpu -q > gpurun_out/r02_gputests_step2_overlap.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_step2_overlap.log); tail -4 gpurun_out/r02_gputests_step2_overlap.log
(timeout 600 python tools/step2_bench.py 200000 65536 > gpurun_out/r02_step2_bench_overlap.txt 2>&1); tail -5 gpurun_out/r02_step2_bench_overlap.txt
(SGB_PROFILE=1 timeout 600 python tools/profile_step1_host.py 200000 500000 > gpurun_out/r02_step1_alloc_profile.txt 2>&1); grep -n "allocation\|native_loops\|wall" gpurun_out/r02_step1_alloc_profile.txt | tail -60
