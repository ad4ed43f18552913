set -x
(timeout 1500 python -m pytest tests/ -m gpu -q -x > gpurun_out/r02_gputests_8.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_8.log); tail -4 gpurun_out/r02_gputests_8.log
(timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu_v6.json 2> gpurun_out/r02_bench_1gpu_v6.err); tail -3 gpurun_out/r02_bench_1gpu_v6.err
python - <<PY
import json
for line in open("gpurun_out/r02_bench_1gpu_v6.json"):
    if line.startswith("{"):
        d=json.loads(line)
        for k in ("value","ms_per_step","e2e","roofline","clocks","ingest","step2","step1"): print(k, str(d.get(k))[:1500])
PY
(timeout 600 python tools/step2_bench.py 200000 65536 > gpurun_out/r02_step2_bench_overlap2.txt 2>&1); tail -5 gpurun_out/r02_step2_bench_overlap2.txt
