set -x
(timeout 900 python -m pytest tests/test_gpu_driver_loops.py -m gpu -x -q > gpurun_out/r02_gputests_loops.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_loops.log); tail -25 gpurun_out/r02_gputests_loops.log
(timeout 900 python bench.py --steps 10 --warmup 3 --no-dense --no-cpu-baseline --no-ingest --no-step2 > gpurun_out/r02_bench_1gpu_loops.json 2> gpurun_out/r02_bench_1gpu_loops.err); tail -5 gpurun_out/r02_bench_1gpu_loops.err
python - <<PY
import json
for line in open("gpurun_out/r02_bench_1gpu_loops.json"):
    if line.startswith("{"):
        d=json.loads(line)
        print(json.dumps(d.get("step1"), indent=1))
PY
