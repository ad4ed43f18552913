set -x
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k two_ranks > gpurun_out/r02_gputests_multi2c.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_multi2c.log); tail -12 gpurun_out/r02_gputests_multi2c.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-dense --no-cpu-baseline --no-ingest --no-step2 > gpurun_out/r02_bench_2gpu_loops.json 2> gpurun_out/r02_bench_2gpu_loops.err); tail -3 gpurun_out/r02_bench_2gpu_loops.err
python - <<PY
import json
for line in open("gpurun_out/r02_bench_2gpu_loops.json"):
    if line.startswith("{"):
        d=json.loads(line)
        s=d.get("step1"); s.pop("driver",None); s.pop("note",None)
        print(d["value"], json.dumps(s))
PY
