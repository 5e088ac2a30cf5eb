cd $GRAFT_REPO_ROOT
one() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], '%.3f' % d['roofline']['frac'], repr(d['checksum_lnl']), d.get('device_front_end',{}).get('value'))"; }
for g in ${GRIDS:-8 32 64 100000}; do for c in cfg4_relbin mb; do BB_RED_GRID_PER_SM=$g timeout 300 python bench_configs.py --config $c 2>/dev/null | one "grid$g:$c"; done; done
