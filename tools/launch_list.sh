# usage: bash tools/launch_list.sh <tag> <bench_configs config> [extra args]  -> gpurun_out/<tag>_launches.csv + per-kernel totals
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tag=$1; cfg=$2; shift 2
timeout 600 ncu ${NCU_FILTER:+-k regex:$NCU_FILTER} --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench_configs.py --config $cfg --steps 2 --warmup 1 "$@" > gpurun_out/${tag}_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/${tag}_launches.csv 2>/dev/null | head -30
