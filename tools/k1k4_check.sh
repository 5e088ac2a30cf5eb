cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg1', d['value'], d['e2e']['value'], d['roofline']['frac'], d['checksum_lnl'])"
for lib in build/libk1_*.so; do BILBY_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['value'], d['e2e']['value'], d['roofline']['frac'], d['checksum_lnl'])"; done
for c in cfg2 cfg0 cfg3; do timeout 300 python bench_configs.py --config $c 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', d['value'], d['e2e']['value'], d['roofline']['frac'], d['checksum_lnl'])"; done
timeout 800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_size.py tests/test_gpu_bns.py tests/test_gpu_reduced_cal.py tests/test_gpu_recon.py -x -q 2>&1 | tail -3
