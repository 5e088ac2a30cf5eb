# usage: bash tools/ncu_full.sh <name> <kernel regex> <skip> <bench_configs config> [extra args]
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
name=$1; pat=$2; skip=$3; cfg=$4; shift 4
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$pat -s $skip -c 1 -f -o gpurun_out/$name python bench_configs.py --config $cfg --steps 1 --warmup 3 "$@" > gpurun_out/$name.log 2>&1
tail -3 gpurun_out/$name.log | cut -c1-300
ls -la gpurun_out/$name.ncu-rep
