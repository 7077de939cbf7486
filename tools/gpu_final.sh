mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12 > gpurun_out/r02z_pytest.log; tail -3 gpurun_out/r02z_pytest.log
( time python bench.py > gpurun_out/r02z_bench1.json 2> gpurun_out/r02z_bench1.err ) 2>&1 | grep real
( time python bench.py --impl reference > gpurun_out/r02z_ref1.json 2> gpurun_out/r02z_ref1.err ) 2>&1 | grep real
python tools/dump_baked_cubin.py cornell_box gpurun_out/r02z_baked.cubin
ncu --set full --import-source on --clock-control none -k regex:k_render_baked -c 1 -f -o gpurun_out/r02z_baked python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic > gpurun_out/r02z_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-traffic > gpurun_out/r02z_launch_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
