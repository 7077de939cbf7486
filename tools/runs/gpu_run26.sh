P="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic --scene stress --width 1920 --height 1080 --spp-per-step 16"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:46])'
for n in 3000 6000 12000 16000 24000 50000 200000; do
  unset ARE_B200_LIB
  echo -n "n=$n default: "; $P --n-prims $n 2>/dev/null | python -c "$S"
  export ARE_B200_LIB=$PWD/variants/libare_b200_big1k.so
  echo -n "n=$n big>1024 quant: "; $P --n-prims $n 2>/dev/null | python -c "$S"
  echo -n "n=$n big>1024 fp32: "; $P --n-prims $n --no-quant 2>/dev/null | python -c "$S"
done
