set -x
nvidia-smi -L
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_gputests_multi2.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_multi2.log); tail -8 gpurun_out/r02_gputests_multi2.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err); tail -3 gpurun_out/r02_bench_2gpu.err
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_ref_2gpu.json 2> gpurun_out/r02_bench_ref_2gpu.err); grep -o '"cores": [0-9]*' gpurun_out/r02_bench_ref_2gpu.json | head -2
python - <<EOF
import json
for line in open("gpurun_out/r02_bench_2gpu.json"):
    if line.startswith("{"):
        d=json.loads(line)
        for k in ("value","ms_per_step","ms_per_step_median","e2e","roofline","ingest","step2","clocks","step1","c4_dense_grm"): print(k, d.get(k))
EOF
