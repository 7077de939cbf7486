P="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["rays_per_sample"],4), r["counted"]["node_visits"], r["counted"]["sphere_tests"])'
for v in default lf ld lfd default; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "rtiow $v: "; $P --scene rtiow_final --width 1200 --height 675 --spp-per-step 100 2>/dev/null | python -c "$S"
done
