mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
python tools/dump_baked_cubin.py textured gpurun_out/r02f_tex_baked.cubin
ncu --set full --import-source on --clock-control none -k regex:k_render_baked -c 1 -f -o gpurun_out/r02f_tex_baked python bench.py --scene textured --width 1920 --height 1080 --spp-per-step 32 $B > gpurun_out/r02f_tex.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_render_path -c 1 -f -o gpurun_out/r02f_rtiow python bench.py --scene rtiow_final --width 1200 --height 675 --spp-per-step 50 $B > gpurun_out/r02f_rtiow.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_render_path -c 1 -f -o gpurun_out/r02f_stress python bench.py --scene stress --width 3840 --height 2160 --spp-per-step 4 $B > gpurun_out/r02f_stress.log 2>&1
ls -la gpurun_out/r02f_*.ncu-rep
