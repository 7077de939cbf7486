P="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["rays_per_sample"],2), "rays/sample", r["kernel"][:40])'
for n in 4 6 8 10 16; do
  echo -n "cloud n=$n auto: "; $P --scene stress --width 1920 --height 1080 --spp-per-step 64 --n-prims $n 2>/dev/null | python -c "$S"
  echo -n "cloud n=$n bvh2: "; $P --scene stress --width 1920 --height 1080 --spp-per-step 64 --n-prims $n --traversal 2 2>/dev/null | python -c "$S"
done
echo -n "textured auto: "; $P --scene textured --width 1920 --height 1080 --spp-per-step 128 2>/dev/null | python -c "$S"
echo -n "textured bvh2: "; $P --scene textured --width 1920 --height 1080 --spp-per-step 128 --traversal 2 2>/dev/null | python -c "$S"
echo -n "cornell auto: "; $P --spp-per-step 64 2>/dev/null | python -c "$S"
echo -n "cornell bvh2: "; $P --spp-per-step 64 --traversal 2 2>/dev/null | python -c "$S"
echo -n "rt_cornell-as-path auto: "; $P --scene rt_cornell_diffuse --width 1024 --height 1024 --spp-per-step 64 2>/dev/null | python -c "$S"
