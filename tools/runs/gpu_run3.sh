mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40 > gpurun_out/r02c_pytest.log; tail -3 gpurun_out/r02c_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"], "frac", round(r["frac"],4))'
echo -n "textured: "; $B --scene textured --width 1920 --height 1080 --spp-per-step 128 2>gpurun_out/r02c_tex.err | python -c "$S"
echo -n "rtiow: "; $B --scene rtiow_final --width 1200 --height 675 --spp-per-step 100 2>/dev/null | python -c "$S"
for P in 0 50 100; do echo -n "stress l2-persist $P (lbvh): "; $B --scene stress --width 3840 --height 2160 --spp-per-step 4 --builder 1 --l2-persist $P 2>gpurun_out/r02c_stress$P.err | python -c "$S"; done
echo -n "stress sah: "; $B --scene stress --width 3840 --height 2160 --spp-per-step 4 --builder 0 2>/dev/null | python -c "$S"
ncu --set full --import-source on --clock-control none -k regex:k_render_path -c 1 -f -o gpurun_out/r02c_textured python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic --scene textured --width 1920 --height 1080 --spp-per-step 16 > gpurun_out/r02c_ncu.log 2>&1
