mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:30], round(r["frac"],4))'
TEX="--scene textured --width 1920 --height 1080"
echo -n "textured baked 256spp: "; $B $TEX --spp-per-step 256 2>/dev/null | python -c "$S"
echo -n "textured baked 128spp: "; $B $TEX --spp-per-step 128 2>/dev/null | python -c "$S"
echo -n "textured generic 128spp: "; $B $TEX --spp-per-step 128 --kernel lean 2>/dev/null | python -c "$S"
python tools/dump_baked_cubin.py textured gpurun_out/r02u_baked.cubin
ncu --set full --import-source on --clock-control none -k regex:k_render_baked -c 1 -f -o gpurun_out/r02u_baked python bench.py $TEX --spp-per-step 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic > gpurun_out/r02u_ncu.log 2>&1
