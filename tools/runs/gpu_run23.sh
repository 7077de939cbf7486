P="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"][:60], round(r["frac"],4))'
for spp in 2 4 8 16 32 64; do
  echo -n "stress-1M spp/step $spp: "; $P --scene stress --width 3840 --height 2160 --spp-per-step $spp 2>/dev/null | python -c "$S"
done
for spp in 16 100 250; do
  echo -n "rtiow spp/step $spp: "; $P --scene rtiow_final --width 1200 --height 675 --spp-per-step $spp 2>/dev/null | python -c "$S"
done
