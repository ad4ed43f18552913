set -x
nvidia-smi -L | wc -l
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_gputests_multi8.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_multi4.log); tail -6 gpurun_out/r02_gputests_multi4.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err); tail -5 gpurun_out/r02_bench_8gpu.err
python - <<EOF
import json
for line in open("gpurun_out/r02_bench_8gpu.json"):
    if line.startswith("{"):
        d=json.loads(line)
        for k in ("value","ms_per_step","ms_per_step_median","e2e","roofline","ingest","step2","clocks","step1","c4_dense_grm","dense_grm"): print(k, d.get(k))
EOF
