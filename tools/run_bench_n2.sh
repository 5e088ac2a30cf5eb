cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err ) 2>&1 | tail -4
tail -5 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n$N.json") if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d["roofline"]["frac"])
for k, v in d["extra"]["configs"].items():
    print(k, v.get("error") or (v["value"], v["e2e"]["value"], round(v["roofline"]["frac"], 3)))
for k, v in d["extra"].get("frequency_sharded_configs3", {}).items():
    print(k, v.get("error") or (v["value"], v["ms_per_step"], v["max_rel_diff_vs_unsharded"], v["exchange_us_per_step_nccl_allreduce_alone"]))
PY
