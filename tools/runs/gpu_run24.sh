P="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms")'
for v in v32_20 v16_24 v16_16 v32_26 v24_20; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "stress-1M 16spp $v: "; $P --scene stress --width 3840 --height 2160 --spp-per-step 16 2>/dev/null | python -c "$S"
  echo -n "stress-1M  4spp $v: "; $P --scene stress --width 3840 --height 2160 --spp-per-step 4 2>/dev/null | python -c "$S"
done
