B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms")'
for mb in 0 5 6 7 9 10; do echo -n "baked min_blocks=$mb: "; $B --baked-min-blocks $mb 2>/dev/null | python -c "$S"; done
