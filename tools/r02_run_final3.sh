set -x
(timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke_final.log 2>&1; echo rc=$? >> gpurun_out/r02_smoke_final.log); tail -3 gpurun_out/r02_smoke_final.log
(timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu_v7.json 2> gpurun_out/r02_bench_1gpu_v7.err); tail -3 gpurun_out/r02_bench_1gpu_v7.err
python - <<PY
import json
for line in open("gpurun_out/r02_bench_1gpu_v7.json"):
    if line.startswith("{"):
        d=json.loads(line)
        print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["whole_product_frac"], d["clocks"], "launches", d.get("gpu_launches"))
        for b in d["batched"]: print(b["k"], b["digits7"]["ms_per_product"], b["digits5"]["ms_per_product"])
        s=d["step1"]; print({k:s.get(k) for k in ("wall_s","first_call_wall_s","variance_ratio_s","wall_with_setgeno_s")}, s["digits5"]["wall_s"], s["r_mirror"]["wall_s"])
        print(d["step2"]["variants_per_s"], d["step2"]["variants_per_s_without_spa"], d["ingest"]["seconds"])
PY
