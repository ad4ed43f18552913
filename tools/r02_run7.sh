set -x
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_4.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_4.log); tail -12 gpurun_out/r02_gputests_4.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-step1 --no-step2 --no-dense > gpurun_out/r02_bench_2gpu_b.json 2> gpurun_out/r02_bench_2gpu_b.err); tail -3 gpurun_out/r02_bench_2gpu_b.err
python - <<EOF
import json
for line in open("gpurun_out/r02_bench_2gpu_b.json"):
    if line.startswith("{"):
        d=json.loads(line)
        for k in ("value","ingest"): print(k, d.get(k))
EOF
