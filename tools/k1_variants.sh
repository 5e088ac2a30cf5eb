cd $GRAFT_REPO_ROOT
for lib in build/libk1_*.so; do
  echo "== $lib"
  BILBY_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['checksum_lnl'])"
done
