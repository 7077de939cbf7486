#!/bin/bash
# tools/bench_scenes.sh — device-resident throughput of the non-headline configs (RTIOW BVH, textured, stress 1M)
P="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"])'
for T in ${TRAVERSALS:-2 3}; do
echo -n "rtiow trav=$T: "; $P --scene rtiow_final --width 1200 --height 675 --spp-per-step 16 --traversal $T 2>/dev/null | python -c "$S"
echo -n "stress-1M trav=$T: "; $P --scene stress --width 3840 --height 2160 --spp-per-step 2 --traversal $T 2>/dev/null | python -c "$S"
done
