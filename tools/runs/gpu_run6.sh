mkdir -p gpurun_out
python -m pytest tests/test_gpu_lbvh.py -q -x -k "bvh4 or refit" 2>&1 | tail -5
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"][:40], "frac", round(r["frac"],4), "nodes/ray", round(r["counted"]["node_visits"]/max(1,r["counted"]["rays"]),1))'
ST="--scene stress --width 3840 --height 2160 --spp-per-step 4 --builder 0"
RT="--scene rtiow_final --width 1200 --height 675 --spp-per-step 100"
echo -n "stress bvh2 (sah): "; $B $ST --traversal 2 2>/dev/null | python -c "$S"
for v in default b4_6 b4_10 b4_12; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "stress bvh4 $v: "; $B $ST --traversal 4 2>gpurun_out/r02g_$v.err | python -c "$S"
done
unset ARE_B200_LIB
echo -n "rtiow bvh2: "; $B $RT --traversal 2 2>/dev/null | python -c "$S"
echo -n "rtiow bvh4: "; $B $RT --traversal 4 2>/dev/null | python -c "$S"
ncu --set full --clock-control none -k regex:k_render_path -c 1 -f -o gpurun_out/r02g_stress_bvh4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic $ST --traversal 4 --spp-per-step 1 > gpurun_out/r02g_ncu.log 2>&1
