#!/bin/bash
# Round-1 second profiling sweep (after the device math layer and the two-kernel time marginalisation):
# one `ncu --set full` capture per kernel; raw metric page + per-instruction SASS page exported on the box.
NCU="ncu --set full --clock-control none --import-source on"
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/$1.sass.csv.gz
    rm -f gpurun_out/$1.ncu-rep
}
$NCU -k regex:bb_inner_product -s 3 -c 1 -o gpurun_out/r1b_k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 200000 > gpurun_out/ncu_k1.log 2>&1
export_rep r1b_k1
$NCU -k regex:bb_series_fill -s 29 -c 1 -o gpurun_out/r1b_k4a python bench_configs.py --config cfg2 --steps 1 > gpurun_out/ncu_k4a.log 2>&1
export_rep r1b_k4a
$NCU -k regex:bb_series_fft -s 29 -c 1 -o gpurun_out/r1b_k4b python bench_configs.py --config cfg2 --steps 1 > gpurun_out/ncu_k4b.log 2>&1
export_rep r1b_k4b
$NCU -k regex:bb_roq_kernel -s 2 -c 1 -o gpurun_out/r1b_k6 python bench_configs.py --config cfg4_roq --batch 200000 --steps 1 > gpurun_out/ncu_k6.log 2>&1
export_rep r1b_k6
$NCU -k regex:bb_relbin -s 2 -c 1 -o gpurun_out/r1b_k5 python bench_configs.py --config cfg4_relbin --batch 200000 --steps 1 > gpurun_out/ncu_k5.log 2>&1
export_rep r1b_k5
ls -la gpurun_out
