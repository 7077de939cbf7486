mkdir -p gpurun_out
( time python bench.py > gpurun_out/r02e_bench1.json 2> gpurun_out/r02e_bench1.err ) 2>&1 | grep real
python tools/dump_baked_cubin.py cornell_box gpurun_out/r02e_baked.cubin
ncu --set full --import-source on --clock-control none -k regex:k_render_baked -c 1 -f -o gpurun_out/r02e_baked python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic > gpurun_out/r02e_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-traffic > gpurun_out/r02e_launch_bench.log 2>&1
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"][:40], "frac", round(r["frac"],4))'
echo -n "textured: "; $B --scene textured --width 1920 --height 1080 --spp-per-step 128 2>/dev/null | python -c "$S"
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
