set -x
nvidia-smi -L | wc -l
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k all_visible > gpurun_out/r02_gputests_multi8b.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_multi8b.log); tail -6 gpurun_out/r02_gputests_multi8b.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_8gpu_b.json 2> gpurun_out/r02_bench_8gpu_b.err); tail -5 gpurun_out/r02_bench_8gpu_b.err
(SGB_PROFILE_RANK0=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 5 --warmup 3 --no-dense --no-cpu-baseline --no-ingest --no-step2 > gpurun_out/r02_bench_8gpu_phases.json 2> gpurun_out/r02_bench_8gpu_phases.err); grep "sgb" gpurun_out/r02_bench_8gpu_phases.err | grep -v "pcg_solve\|K.\[" | head -60
python - <<PY
import json
for line in open("gpurun_out/r02_bench_8gpu_b.json"):
    if line.startswith("{"):
        d=json.loads(line)
        for k in ("value","ms_per_step","ms_per_step_median","e2e","roofline","ingest","step2","clocks","c4_dense_grm","dense_grm"): print(k, d.get(k))
        s=d.get("step1"); s.pop("driver",None); s.pop("note",None); print("step1", json.dumps(s))
PY
