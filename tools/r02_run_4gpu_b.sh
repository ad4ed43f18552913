set -x
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02_bench_4gpu_b.json 2> gpurun_out/r02_bench_4gpu_b.err); tail -3 gpurun_out/r02_bench_4gpu_b.err
python - <<PY
import json
for line in open("gpurun_out/r02_bench_4gpu_b.json"):
    if line.startswith("{"):
        d=json.loads(line)
        for k in ("value","ms_per_step","ms_per_step_median","e2e","roofline","ingest","step2","clocks","c4_dense_grm"): print(k, str(d.get(k))[:900])
        s=d.get("step1"); s.pop("driver",None); s.pop("note",None); print("step1", json.dumps(s))
PY
