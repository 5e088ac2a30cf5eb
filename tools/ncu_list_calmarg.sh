cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches_calmarg.csv python bench_configs.py --config calmarg --steps 1 --warmup 3 > gpurun_out/r2_ncu_calmarg.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_calmarg.csv 2>&1 | tail -12
tail -1 gpurun_out/r2_ncu_calmarg.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['algorithmic_flop_per_step'], d['roofline']['units_per_eval'])"
