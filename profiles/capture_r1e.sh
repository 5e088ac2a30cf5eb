#!/bin/bash
# Round-1 closing captures (after K0 staging, butterfly reductions, time + calibration kernel)
NCU="ncu --set full --clock-control none --import-source on"
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/$1.sass.csv.gz
    rm -f gpurun_out/$1.ncu-rep
}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_v5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
$NCU -k regex:bb_prologue -s 4 -c 1 -o gpurun_out/r1e_k0 python bench_configs.py --config cfg4_relbin --batch 1000000 --steps 1 > gpurun_out/ncu_k0.log 2>&1
export_rep r1e_k0
$NCU -k regex:bb_relbin -s 2 -c 1 -o gpurun_out/r1e_k5 python bench_configs.py --config cfg4_relbin --batch 200000 --steps 1 > gpurun_out/ncu_k5.log 2>&1
export_rep r1e_k5
$NCU -k regex:bb_roq_kernel -s 2 -c 1 -o gpurun_out/r1e_k6 python bench_configs.py --config cfg4_roq --batch 200000 --steps 1 > gpurun_out/ncu_k6.log 2>&1
export_rep r1e_k6
$NCU -k regex:bb_calmarg_time -s 2 -c 1 -o gpurun_out/r1e_kct python bench_configs.py --config calmarg_time --batch 256 --steps 1 > gpurun_out/ncu_kct.log 2>&1
export_rep r1e_kct
