#!/bin/bash
# Round-1 third profiling sweep (K5 edge form, K6 row combination): one `ncu --set full` capture per kernel.
NCU="ncu --set full --clock-control none --import-source on"
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/$1.sass.csv.gz
    rm -f gpurun_out/$1.ncu-rep
}
$NCU -k regex:bb_roq_kernel -s 2 -c 1 -o gpurun_out/r1c_k6 python bench_configs.py --config cfg4_roq --batch 200000 --steps 1 > gpurun_out/ncu_k6.log 2>&1
export_rep r1c_k6
$NCU -k regex:bb_relbin -s 2 -c 1 -o gpurun_out/r1c_k5 python bench_configs.py --config cfg4_relbin --batch 200000 --steps 1 > gpurun_out/ncu_k5.log 2>&1
export_rep r1c_k5
$NCU -k regex:bb_prologue -s 2 -c 1 -o gpurun_out/r1c_k0 python bench_configs.py --config cfg4_relbin --batch 200000 --steps 1 > gpurun_out/ncu_k0.log 2>&1
export_rep r1c_k0
ls -la gpurun_out
