set -x
(timeout 900 python -m pytest tests/test_gpu_vs_reference_solver.py -m gpu -q > gpurun_out/r02_gputests_refsolver.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_refsolver.log); tail -8 gpurun_out/r02_gputests_refsolver.log
(timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err); tail -2 gpurun_out/r02_bench_reference_arm.err; cut -c1-900 gpurun_out/r02_bench_reference_arm.json
(timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu_v5.json 2> gpurun_out/r02_bench_1gpu_v5.err); tail -3 gpurun_out/r02_bench_1gpu_v5.err
python - <<PY
import json
for line in open("gpurun_out/r02_bench_1gpu_v5.json"):
    if line.startswith("{"):
        d=json.loads(line)
        for k in ("value","ms_per_step","e2e","roofline","clocks","cpu_baseline"): print(k, str(d.get(k))[:900])
PY
