#!/bin/bash
# tools/ab_scenes.sh v1 v2 ... — tools/bench_scenes.sh with each variants/libare_b200_<v>.so ("default" = the in-tree library)
for v in "$@"; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo "== $v"; bash tools/bench_scenes.sh
done
