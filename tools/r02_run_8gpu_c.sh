set -x
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 3 --no-ingest --no-step2 --no-cpu-baseline > gpurun_out/r02_bench_8gpu_c.json 2> gpurun_out/r02_bench_8gpu_c.err); tail -3 gpurun_out/r02_bench_8gpu_c.err
python - <<PY
import json
for line in open("gpurun_out/r02_bench_8gpu_c.json"):
    if line.startswith("{"):
        d=json.loads(line)
        for k in ("value","ms_per_step","ms_per_step_median","e2e","clocks","c4_dense_grm","batched"): print(k, str(d.get(k))[:1200])
        s=d.get("step1"); s.pop("driver",None); s.pop("note",None); print("step1", json.dumps(s))
PY
