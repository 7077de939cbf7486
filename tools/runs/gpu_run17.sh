B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:30], round(r["frac"],4))'
TEX="--scene textured --width 1920 --height 1080 --spp-per-step 256"
for mb in 2 3; do echo -n "textured baked mb$mb: "; $B $TEX --baked-min-blocks $mb 2>/dev/null | python -c "$S"; done
for mb in 3 4; do echo -n "cornell baked mb$mb: "; $B --baked-min-blocks $mb 2>/dev/null | python -c "$S"; done
echo -n "textured precompiled generic (default 7): "; $B $TEX --kernel lean 2>/dev/null | python -c "$S"
for v in mb4 mb5; do echo -n "textured precompiled generic $v: "; ARE_B200_LIB=$PWD/variants/libare_b200_$v.so $B $TEX --kernel lean 2>/dev/null | python -c "$S"; done
