P="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic --scene stress --width 1920 --height 1080 --spp-per-step 64"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:50])'
for n in 12 20 28 32 40 64; do
  echo -n "n=$n auto: "; $P --n-prims $n 2>/dev/null | python -c "$S"
  echo -n "n=$n no bake: "; $P --n-prims $n --kernel lean 2>/dev/null | python -c "$S"
  echo -n "n=$n bvh2: "; $P --n-prims $n --traversal 2 2>/dev/null | python -c "$S"
done
