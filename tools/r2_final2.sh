# final round-2 pass on one B200: GPU tests, the driver's bench line (both arms), launch list, ncu captures
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2k_gpu_tests.log; cat gpurun_out/r2k_gpu_tests.log
( time timeout 1400 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2k_bench_n1.err
( time timeout 600 python bench.py --impl reference > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err ) 2>&1 | tail -4
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2k_bench_n1.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["peak"], d["cpu_baseline"]["value"], d["clocks"])
for k, v in d["extra"]["configs"].items():
    print(k, v.get("error") or (v["value"], v["e2e"]["value"], round(v["roofline"]["frac"], 3), v["roofline"]["peak"],
                                v.get("cpu_baseline", {}).get("value"), v.get("device_front_end", {}).get("value")))
print(open("gpurun_out/r2k_bench_ref.json").read()[:400])
PY
bash profiles/capture_r2.sh r2k launches 2>&1 | tail -3
timeout 300 python bench_configs.py --config mb 2>/dev/null | tail -1 > gpurun_out/r2k_bench_mb.json
timeout 300 python bench_configs.py --config calmarg 2>/dev/null | tail -1 > gpurun_out/r2k_bench_calmarg.json
cut -c1-300 gpurun_out/r2k_bench_mb.json gpurun_out/r2k_bench_calmarg.json
