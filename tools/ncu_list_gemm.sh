set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in cfg4_roq_time calmarg; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches_$cfg.csv python bench_configs.py --config $cfg --steps 2 --warmup 3 > gpurun_out/r2_ncu_$cfg.log 2>&1
  tail -2 gpurun_out/r2_ncu_$cfg.log | cut -c1-300
done
python profiles/summarize_launches.py gpurun_out/r2_launches_cfg4_roq_time.csv 2>&1 | tail -15
python profiles/summarize_launches.py gpurun_out/r2_launches_calmarg.csv 2>&1 | tail -15
