set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python bench_configs.py --config cfg4_roq_time 2>&1 | tail -1 > gpurun_out/r2_cfg4_roq_time_$i.json; done
timeout 300 python bench_configs.py --config calmarg 2>&1 | tail -1 > gpurun_out/r2_calmarg.json
python -c "
import json
for f in ('r2_cfg4_roq_time_1','r2_cfg4_roq_time_2','r2_calmarg'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['roofline']['achieved'], d['checksum_lnl'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches_cfg4_roq_time.csv python bench_configs.py --config cfg4_roq_time --steps 2 --warmup 3 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_cfg4_roq_time.csv | tail -6
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bb_gemm_nt -s 2 -c 1 -f -o gpurun_out/r2_gemm_k7 python bench_configs.py --config cfg4_roq_time --steps 1 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
