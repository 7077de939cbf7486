#!/bin/bash
# tools/bench_all_configs.sh — every BASELINE.json configuration on one GPU, each with its CPU baseline timed on the
# same box (bench.py's cpu_baseline leg: the fp64 CPU twin on all host threads), plus for config 0 the REAL reference
# program (oracle/_ref/rt_ref = experiments/rt.cpp unmodified, single-threaded) and its ray count (rt_ref_counted).
# One JSON line per configuration on stdout.
cd "$(dirname "$0")/.."
P="python bench.py --steps 4 --warmup 3 --cpu-seconds 8"
$P --scene rt_cornell --width 512 --height 512 --spp-per-step 1 2>/dev/null
$P --scene rtiow_final --width 1200 --height 675 --spp-per-step 250 2>/dev/null
$P --scene textured --width 1920 --height 1080 --spp-per-step 256 2>/dev/null
$P --scene cornell_box --width 2048 --height 2048 --spp-per-step 256 2>/dev/null
ARE_CUDA_VERBOSE=1 $P --scene stress --width 3840 --height 2160 --spp-per-step 2 2>gpurun_out/stress_commit.err
if [ -x oracle/_ref/rt_ref ]; then
  d=$(mktemp -d); ( cd $d; s=$(date +%s.%N); $OLDPWD/oracle/_ref/rt_ref > rt_ref.log 2>&1; e=$(date +%s.%N); ARE_RT_SEED=1 $OLDPWD/oracle/_ref/rt_ref_counted > cnt.log 2>&1
    python - "$s" "$e" <<'PY'
import json, re, sys
s, e = float(sys.argv[1]), float(sys.argv[2])
m = re.search(r"ARE_COUNT rays=(\d+) tri_tests=(\d+)", open("cnt.log").read())
rays = int(m.group(1)) if m else None
print(json.dumps({"reference_program": "experiments/rt.cpp unmodified (oracle/_ref/rt_ref), 512x512, 1 thread", "wall_s": e - s,
                  "Msamples_per_s": 512 * 512 / (e - s) / 1e6, "rays": rays, "Mrays_per_s": rays / (e - s) / 1e6 if rays else None}))
PY
  ); rm -rf $d
fi
