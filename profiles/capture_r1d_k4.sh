NCU="ncu --set full --clock-control none --import-source on"
export_rep() {
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/$1.sass.csv.gz
    rm -f gpurun_out/$1.ncu-rep
}
$NCU -k regex:bb_series_fill -s 29 -c 1 -o gpurun_out/r1d_k4a python bench_configs.py --config cfg2 --steps 1 > gpurun_out/ncu_k4a.log 2>&1
export_rep r1d_k4a
$NCU -k regex:bb_series_fft -s 29 -c 1 -o gpurun_out/r1d_k4b python bench_configs.py --config cfg2 --steps 1 > gpurun_out/ncu_k4b.log 2>&1
export_rep r1d_k4b
