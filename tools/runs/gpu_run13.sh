python -m pytest tests -m gpu -q 2>&1 | tail -8
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:30], round(r["frac"],4))'
TEX="--scene textured --width 1920 --height 1080 --spp-per-step 128"
echo -n "textured baked: "; $B $TEX 2>/dev/null | python -c "$S"
echo -n "textured generic: "; $B $TEX --kernel lean 2>/dev/null | python -c "$S"
for mb in 5 6 8; do echo -n "textured baked mb$mb: "; $B $TEX --baked-min-blocks $mb 2>/dev/null | python -c "$S"; done
