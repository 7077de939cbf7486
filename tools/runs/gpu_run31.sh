python -m pytest tests -m gpu -q -x 2>&1 | tail -3
P="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:46])'
echo -n "cornell: "; $P 2>/dev/null | python -c "$S"
echo -n "textured: "; $P --scene textured --width 1920 --height 1080 --spp-per-step 256 2>/dev/null | python -c "$S"
echo -n "textured precompiled: "; $P --scene textured --width 1920 --height 1080 --spp-per-step 256 --kernel lean 2>/dev/null | python -c "$S"
echo -n "rtiow: "; $P --scene rtiow_final --width 1200 --height 675 --spp-per-step 100 2>/dev/null | python -c "$S"
echo -n "stress: "; $P --scene stress --width 3840 --height 2160 --spp-per-step 16 2>/dev/null | python -c "$S"
