python -m pytest tests -m gpu -q -x 2>&1 | tail -3
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:30], round(r["frac"],4))'
TEX="--scene textured --width 1920 --height 1080"
for mb in 4 5 6 7; do echo -n "textured baked 256spp mb$mb: "; $B $TEX --spp-per-step 256 --baked-min-blocks $mb 2>/dev/null | python -c "$S"; done
echo -n "textured baked 128spp: "; $B $TEX --spp-per-step 128 2>/dev/null | python -c "$S"
