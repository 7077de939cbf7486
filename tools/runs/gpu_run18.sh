P="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"][:40])'
for v in default bvh4 bvh5 bvh6; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "rtiow $v: "; $P --scene rtiow_final --width 1200 --height 675 --spp-per-step 100 2>/dev/null | python -c "$S"
done
for v in default big8 big10 big16; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "stress-1M $v: "; $P --scene stress --width 3840 --height 2160 --spp-per-step 4 2>/dev/null | python -c "$S"
done
