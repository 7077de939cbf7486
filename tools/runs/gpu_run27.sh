P="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic --scene stress"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:46])'
for v in default le2 le4 le8; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "1M $v: "; $P --width 3840 --height 2160 --spp-per-step 16 2>/dev/null | python -c "$S"
  echo -n "50k $v: "; $P --width 1920 --height 1080 --spp-per-step 16 --n-prims 50000 2>/dev/null | python -c "$S"
  echo -n "12k $v: "; $P --width 1920 --height 1080 --spp-per-step 16 --n-prims 12000 2>/dev/null | python -c "$S"
done
