ncu --set full --import-source on --clock-control none -k regex:k_render_path -c 1 -f -o gpurun_out/r02y_rtiow python bench.py --scene rtiow_final --width 1200 --height 675 --spp-per-step 50 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic > gpurun_out/r02y_rtiow_ncu.log 2>&1
tail -1 gpurun_out/r02y_rtiow_ncu.log | cut -c1-200
