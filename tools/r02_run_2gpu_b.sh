set -x
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_gputests_multi2b.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests_multi2b.log); tail -12 gpurun_out/r02_gputests_multi2b.log
(timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1; echo rc=$? >> gpurun_out/r02_sanitizer_memcheck_smoke.log); tail -6 gpurun_out/r02_sanitizer_memcheck_smoke.log
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "crossprod_matches_oracle and (tensor or umma) and (1 or 3 or 31)" > gpurun_out/r02_sanitizer_memcheck_products.log 2>&1; echo rc=$? >> gpurun_out/r02_sanitizer_memcheck_products.log); tail -6 gpurun_out/r02_sanitizer_memcheck_products.log
(timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r02_sanitizer_racecheck_smoke.log 2>&1; echo rc=$? >> gpurun_out/r02_sanitizer_racecheck_smoke.log); tail -8 gpurun_out/r02_sanitizer_racecheck_smoke.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --c4-full --no-step2 --no-cpu-baseline > gpurun_out/r02_bench_2gpu_c4.json 2> gpurun_out/r02_bench_2gpu_c4.err); tail -3 gpurun_out/r02_bench_2gpu_c4.err
python - <<EOF
import json
for line in open("gpurun_out/r02_bench_2gpu_c4.json"):
    if line.startswith("{"):
        d=json.loads(line)
        for k in ("value","ms_per_step_list","clocks","ingest","c4_dense_grm"): print(k, d.get(k))
EOF
