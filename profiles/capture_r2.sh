#!/bin/bash
# Round-2 captures: launch list of the headline bench command + one `ncu --set full` capture per hot kernel.
# usage (on the GPU box): bash profiles/capture_r2.sh <tag> [kernels...]   -> gpurun_out/<tag>_*.{raw.csv,sass.csv.gz}
cd $GRAFT_REPO_ROOT
tag=$1; shift
NCU="ncu --set full --clock-control none --import-source on"
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/$1.sass.csv.gz
    rm -f gpurun_out/$1.ncu-rep
}
for k in "$@"; do
case $k in
launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_launches.log 2>&1 ;;
k1) timeout 600 $NCU -k regex:bb_inner_product -s 2 -c 1 -f -o gpurun_out/${tag}_k1 python bench_configs.py --config cfg1 --batch 200000 --steps 1 > gpurun_out/${tag}_k1.log 2>&1; export_rep ${tag}_k1 ;;
k1tf2) timeout 600 $NCU -k regex:bb_inner_product -s 2 -c 1 -f -o gpurun_out/${tag}_k1tf2 python bench_configs.py --config cfg3 --batch 2048 --steps 1 > gpurun_out/${tag}_k1tf2.log 2>&1; export_rep ${tag}_k1tf2 ;;
k4a) timeout 600 $NCU -k regex:bb_series_fill -s 4 -c 1 -f -o gpurun_out/${tag}_k4a python bench_configs.py --config cfg2 --batch 100000 --steps 1 > gpurun_out/${tag}_k4a.log 2>&1; export_rep ${tag}_k4a ;;
k4b) timeout 600 $NCU -k regex:bb_series_fft -s 4 -c 1 -f -o gpurun_out/${tag}_k4b python bench_configs.py --config cfg2 --batch 100000 --steps 1 > gpurun_out/${tag}_k4b.log 2>&1; export_rep ${tag}_k4b ;;
k5) timeout 600 $NCU -k regex:bb_relbin -s 2 -c 1 -f -o gpurun_out/${tag}_k5 python bench_configs.py --config cfg4_relbin --batch 200000 --steps 1 > gpurun_out/${tag}_k5.log 2>&1; export_rep ${tag}_k5 ;;
k6) timeout 600 $NCU -k regex:bb_roq_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_k6 python bench_configs.py --config cfg4_roq --batch 200000 --steps 1 > gpurun_out/${tag}_k6.log 2>&1; export_rep ${tag}_k6 ;;
k5mb) timeout 600 $NCU -k regex:bb_relbin -s 2 -c 1 -f -o gpurun_out/${tag}_k5mb python bench_configs.py --config mb --batch 32768 --steps 1 > gpurun_out/${tag}_k5mb.log 2>&1; export_rep ${tag}_k5mb ;;
k0tf2) timeout 600 $NCU -k regex:bb_prologue -s 4 -c 1 -f -o gpurun_out/${tag}_k0tf2 python bench_configs.py --config cfg4_relbin --batch 1000000 --steps 1 > gpurun_out/${tag}_k0tf2.log 2>&1; export_rep ${tag}_k0tf2 ;;
k0) timeout 600 $NCU -k regex:bb_prologue -s 2 -c 1 -f -o gpurun_out/${tag}_k0 python bench_configs.py --config cfg1 --batch 1000000 --steps 1 > gpurun_out/${tag}_k0.log 2>&1; export_rep ${tag}_k0 ;;
esac
done
ls -la gpurun_out/${tag}_* | cut -c30-
