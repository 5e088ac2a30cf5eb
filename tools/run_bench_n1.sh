cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1100 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n1.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["peak"], d["cpu_baseline"]["value"])
for k, v in d["extra"]["configs"].items():
    print(k, v.get("error") or (v["value"], v["e2e"]["value"], round(v["roofline"]["frac"], 3), v["roofline"]["peak"],
                                v.get("cpu_baseline", {}).get("value")))
PY
bash tools/ncu_full.sh r2_k1 bb_inner_product 2 cfg1 --batch 200000 | tail -2
