set -x
(timeout 72 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 3 --no-dense > gpurun_out/r02_bench_8gpu_final.json 2> gpurun_out/r02_bench_8gpu_final.err); tail -3 gpurun_out/r02_bench_8gpu_final.err
python - <<PY
import json
for line in open("gpurun_out/r02_bench_8gpu_final.json"):
    if line.startswith("{"):
        d=json.loads(line)
        print("value", d["value"], "median ms", d.get("ms_per_step_median"), "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["whole_product_frac"], d["clocks"])
        for b in d["batched"]: print(b["k"], b["digits7"]["ms_per_product"], b["digits5"]["ms_per_product"])
        s=d["step1"]; print({k:s.get(k) for k in ("wall_s","first_call_wall_s","variance_ratio_s","wall_with_setgeno_s")}, s["digits5"]["wall_s"], s["r_mirror"]["wall_s"])
        print(d["step2"]["variants_per_s"], d["step2"]["variants_per_s_without_spa"], d["ingest"]["seconds"])
PY
