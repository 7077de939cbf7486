# textured scene through the generic baked kernel: full ncu capture + the CUBIN the library compiled on this box
mkdir -p gpurun_out
python tools/dump_baked_cubin.py textured gpurun_out/r02t_baked.cubin
ncu --set full --import-source on --clock-control none -k regex:k_render_baked -c 1 -f -o gpurun_out/r02t_baked python bench.py --scene textured --width 1920 --height 1080 --spp-per-step 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic > gpurun_out/r02t_ncu.log 2>&1
tail -2 gpurun_out/r02t_ncu.log | cut -c1-300
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", r["kernel"][:30], round(r["frac"],4))'
echo -n "textured baked 256spp: "; $B --scene textured --width 1920 --height 1080 --spp-per-step 256 2>/dev/null | python -c "$S"
echo -n "cornell: "; $B 2>/dev/null | python -c "$S"
