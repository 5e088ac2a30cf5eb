#!/bin/bash
# Round-1 final profiling sweep (K1 with the push argument, K5 edge form, multi-banding, K6 with bulk copies):
# launch list of the headline bench + one `ncu --set full` capture per changed kernel.
NCU="ncu --set full --clock-control none --import-source on"
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/$1.sass.csv.gz
    rm -f gpurun_out/$1.ncu-rep
}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_v4.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
$NCU -k regex:bb_inner_product -s 3 -c 1 -o gpurun_out/r1d_k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 200000 > gpurun_out/ncu_k1.log 2>&1
export_rep r1d_k1
$NCU -k regex:bb_roq_kernel -s 2 -c 1 -o gpurun_out/r1d_k6 python bench_configs.py --config cfg4_roq --batch 200000 --steps 1 > gpurun_out/ncu_k6.log 2>&1
export_rep r1d_k6
$NCU -k regex:bb_relbin -s 2 -c 1 -o gpurun_out/r1d_k5mb python bench_configs.py --config mb --batch 8192 --steps 1 > gpurun_out/ncu_k5mb.log 2>&1
export_rep r1d_k5mb
$NCU -k regex:bb_prologue -c 8 -o gpurun_out/r1d_k0 python bench_configs.py --config cfg4_relbin --batch 1000000 --steps 1 > gpurun_out/ncu_k0.log 2>&1
export_rep r1d_k0
