B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms")'
ST="--scene stress --width 3840 --height 2160 --spp-per-step 4 --builder 1"
RT="--scene rtiow_final --width 1200 --height 675 --spp-per-step 100"
for v in default le2 le4 le2s2 le3s6 le4s8; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "$v rtiow: "; $B $RT 2>/dev/null | python -c "$S"
  echo -n "$v stress: "; $B $ST 2>/dev/null | python -c "$S"
done
