mkdir -p gpurun_out
python -m pytest tests/test_gpu_lbvh.py -q -x -k "bvh4" 2>&1 | tail -3
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"][:40], "frac", round(r["frac"],4), "nodes/ray", round(r["counted"]["node_visits"]/max(1,r["counted"]["rays"]),1))'
ST="--scene stress --width 3840 --height 2160 --spp-per-step 4 --builder 0"
for v in b4_6 b4_10 b4_12; do
  export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so
  echo -n "stress bvh4 $v: "; $B $ST --traversal 4 2>gpurun_out/r02g_$v.err | python -c "$S"
done
