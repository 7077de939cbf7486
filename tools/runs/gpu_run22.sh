P="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"][:60])'
for v in default ml3 ml12 ml20 ml28 spv2 spv8; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "stress-1M $v: "; $P --scene stress --width 3840 --height 2160 --spp-per-step 4 2>/dev/null | python -c "$S"
  echo -n "rtiow $v: "; $P --scene rtiow_final --width 1200 --height 675 --spp-per-step 100 2>/dev/null | python -c "$S"
done
unset ARE_B200_LIB
ncu --set full --import-source on --clock-control none -k regex:k_render_path -c 1 -f -o gpurun_out/r02y_stress_quant python bench.py --scene stress --width 3840 --height 2160 --spp-per-step 2 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic > gpurun_out/r02y_ncu.log 2>&1
tail -1 gpurun_out/r02y_ncu.log | cut -c1-200
