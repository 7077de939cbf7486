python -m pytest tests/test_gpu_pool.py -q -x 2>&1 | tail -5
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:30])'
ST="--scene stress --width 3840 --height 2160 --spp-per-step 4 --builder 1"
RT="--scene rtiow_final --width 1200 --height 675 --spp-per-step 100"
echo -n "megakernel rtiow: "; $B $RT 2>/dev/null | python -c "$S"
echo -n "megakernel stress: "; $B $ST 2>/dev/null | python -c "$S"
for v in default p32 p96 p128 pmb7 pmb6; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "pool $v rtiow: "; $B $RT --pool --traversal 2 2>gpurun_out/r02h_$v.err | python -c "$S"
  echo -n "pool $v stress: "; $B $ST --pool --traversal 2 2>/dev/null | python -c "$S"
done
unset ARE_B200_LIB
echo -n "pool cornell-bvh 1024: "; $B --scene cornell_box --width 1024 --height 1024 --spp-per-step 64 --pool --traversal 2 2>/dev/null | python -c "$S"
echo -n "pool textured-bvh: "; $B --scene textured --width 1920 --height 1080 --spp-per-step 128 --pool --traversal 2 2>/dev/null | python -c "$S"
echo -n "mega textured-bvh: "; $B --scene textured --width 1920 --height 1080 --spp-per-step 128 --traversal 2 2>/dev/null | python -c "$S"
ncu --set full --clock-control none -k regex:k_render_pool -c 1 -f -o gpurun_out/r02h_pool_rtiow python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic $RT --spp-per-step 16 --pool --traversal 2 > gpurun_out/r02h_ncu.log 2>&1
