#!/bin/bash
# Round-1 profiling sweep (run under gpurun on one B200): launch list of the headline bench + one
# `ncu --set full` capture per kernel.  Reports land in gpurun_out/; profiles/summarize_ncu.py turns them into the
# committed summaries.
set -x
NCU="ncu --set full --clock-control none --import-source on"
# the reports are 25-30 MB each and gpurun brings back at most 64 MiB: export the raw metrics page (and the
# per-instruction SASS page, gzipped) on the box and drop the report
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/$1.sass.csv.gz
    rm -f gpurun_out/$1.ncu-rep
}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_v3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
$NCU -k regex:"bb_inner_product|bb_prologue|bb_epilogue" -s 6 -c 3 -o gpurun_out/r1_k013_v3 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 200000 > gpurun_out/ncu_k013.log 2>&1
export_rep r1_k013_v3
$NCU -k regex:bb_inner_product -s 2 -c 1 -o gpurun_out/r1_k1_taylorf2 \
    python bench_configs.py --config cfg3 --batch 2048 --steps 1 > gpurun_out/ncu_k1tf2.log 2>&1
export_rep r1_k1_taylorf2
$NCU -k regex:bb_time_marg -s 2 -c 1 -o gpurun_out/r1_k4_v3 \
    python bench_configs.py --config cfg2 --batch 20000 --steps 1 > gpurun_out/ncu_k4.log 2>&1
export_rep r1_k4_v3
$NCU -k regex:bb_relbin -s 2 -c 1 -o gpurun_out/r1_k5 \
    python bench_configs.py --config cfg4_relbin --batch 200000 --steps 1 > gpurun_out/ncu_k5.log 2>&1
export_rep r1_k5
$NCU -k regex:bb_roq_kernel -s 2 -c 1 -o gpurun_out/r1_k6 \
    python bench_configs.py --config cfg4_roq --batch 200000 --steps 1 > gpurun_out/ncu_k6.log 2>&1
export_rep r1_k6
$NCU -k regex:"bb_roq_hlinear|bb_roq_time_marg|gemm" -s 10 -c 5 -o gpurun_out/r1_k7 \
    python bench_configs.py --config cfg4_roq_time --batch 8192 --steps 1 > gpurun_out/ncu_k7.log 2>&1
export_rep r1_k7
for f in gpurun_out/ncu_k*.log; do tail -n 2 $f; done
