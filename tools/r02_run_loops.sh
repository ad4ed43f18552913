set -x
(SGB_PROFILE=1 timeout 600 python tools/profile_step1_host.py 200000 62500 > gpurun_out/r02_step1_phases_M62500.txt 2>&1); grep -v "K.\[PY" gpurun_out/r02_step1_phases_M62500.txt | tail -150
