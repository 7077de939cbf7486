python -m pytest tests -m gpu -q -x 2>&1 | tail -3
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:30], round(r["frac"],4))'
echo -n "cornell: "; $B 2>/dev/null | python -c "$S"
echo -n "cornell mb5: "; $B --baked-min-blocks 5 2>/dev/null | python -c "$S"
echo -n "cornell lean: "; $B --kernel lean 2>/dev/null | python -c "$S"
echo -n "rtiow: "; $B --scene rtiow_final --width 1200 --height 675 --spp-per-step 100 2>/dev/null | python -c "$S"
echo -n "textured: "; $B --scene textured --width 1920 --height 1080 --spp-per-step 128 2>/dev/null | python -c "$S"
echo -n "stress: "; $B --scene stress --width 3840 --height 2160 --spp-per-step 4 --builder 1 2>/dev/null | python -c "$S"
