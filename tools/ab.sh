#!/bin/bash
# tools/ab.sh v1 v2 ... — time bench.py with each variants/libare_b200_<v>.so ("default" = the in-tree library)
for v in "$@"; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "== $v: "
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $AB_ARGS 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],3), 'ms', round(d['mrays_per_s'],1), 'Mrays/s')"
done
