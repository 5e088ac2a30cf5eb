# e2e of the headline configuration for several host-chunk settings (rows): "<chunk> <first>" ...
cd $GRAFT_REPO_ROOT
one() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '%.4g' % d['value'], '%.4g' % d['e2e']['value'])"; }
for pair in "$@"; do set -- $pair; BB_HOST_CHUNK=$1 BB_HOST_FIRST_CHUNK=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline | one "chunk=$1,first=$2"; done
