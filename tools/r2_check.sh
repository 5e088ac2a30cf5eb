# quick check of a kernel change on the GPU box: headline + per-configuration benches, then the parity tests
cd $GRAFT_REPO_ROOT
one() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], '%.3f' % d['roofline']['frac'], repr(d['checksum_lnl']))"; }
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline | one cfg1
for lib in build/libvar_*.so; do [ -f "$lib" ] && BILBY_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline | one $lib; done
for c in ${CONFIGS:-cfg2 cfg0 cfg3 cfg4_relbin cfg4_roq mb}; do timeout 300 python bench_configs.py --config $c 2>/dev/null | one $c; done
[ -n "$SKIP_TESTS" ] || timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
