# usage: bash tools/r2_variants.sh "<lib> <config>" ...   (bench_configs lines for experiment libraries)
cd $GRAFT_REPO_ROOT
one() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], '%.3f' % d['roofline']['frac'], repr(d['checksum_lnl']))"; }
for pair in "$@"; do set -- $pair; BILBY_B200_LIB=$PWD/$1 timeout 300 python bench_configs.py --config $2 2>/dev/null | one "$1:$2"; done
