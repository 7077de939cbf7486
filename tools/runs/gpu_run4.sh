mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 -x 2>&1 | tail -30 > gpurun_out/r02d_pytest.log; tail -3 gpurun_out/r02d_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"][:40], "frac", round(r["frac"],4))'
TEX="--scene textured --width 1920 --height 1080 --spp-per-step 128"
RT="--scene rtiow_final --width 1200 --height 675 --spp-per-step 100"
ST="--scene stress --width 3840 --height 2160 --spp-per-step 4 --builder 1"
for v in default mb7 mb8 ss8 ss16; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo "== $v"
  if [ "$v" = default -o "$v" = mb7 -o "$v" = mb8 ]; then echo -n "  textured: "; $B $TEX 2>/dev/null | python -c "$S"; fi
  echo -n "  rtiow: "; $B $RT 2>/dev/null | python -c "$S"
  if [ "$v" != mb7 -a "$v" != mb8 ]; then echo -n "  stress: "; $B $ST 2>/dev/null | python -c "$S"; fi
done
unset ARE_B200_LIB
echo "== wavefront"
echo -n "  rtiow: "; $B $RT --wavefront --traversal 2 2>gpurun_out/r02d_wf.err | python -c "$S"
echo -n "  cornell 1024^2 64spp: "; $B --scene cornell_box --width 1024 --height 1024 --spp-per-step 64 --wavefront --traversal 2 2>/dev/null | python -c "$S"
echo -n "  cornell 1024^2 64spp megakernel bvh2: "; $B --scene cornell_box --width 1024 --height 1024 --spp-per-step 64 --traversal 2 2>/dev/null | python -c "$S"
echo -n "  stress: "; $B $ST --wavefront --traversal 2 2>/dev/null | python -c "$S"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_wf --csv --log-file gpurun_out/r02d_wf_kernels.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic $RT --spp-per-step 16 --wavefront --traversal 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_render_path -c 3 --csv --log-file gpurun_out/r02d_mega_kernels.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic $RT --spp-per-step 16 > /dev/null 2>&1
