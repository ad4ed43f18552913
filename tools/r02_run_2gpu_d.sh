set -x
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k two_ranks > gpurun_out/r02_gputests_multi2d.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_multi2d.log); tail -6 gpurun_out/r02_gputests_multi2d.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_2gpu_final.json 2> gpurun_out/r02_bench_2gpu_final.err); tail -3 gpurun_out/r02_bench_2gpu_final.err
python - <<PY
import json
for line in open("gpurun_out/r02_bench_2gpu_final.json"):
    if line.startswith("{"):
        d=json.loads(line)
        for k in ("value","ms_per_step","e2e","roofline","clocks","ingest","step2"): print(k, str(d.get(k))[:700])
        s=d.get("step1") or {}; s.pop("driver",None); s.pop("note",None); print("step1", str(s)[:900])
PY
