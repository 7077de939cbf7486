python -m pytest tests/test_gpu_lbvh.py tests/test_gpu_fullsize.py -m gpu -q -x -s 2>&1 | grep -E "quantised|stress 1M|passed|failed|Error|assert" | head -20
P="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"][:60])'
for v in default q8 q10 q16 qsc2; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "stress-1M $v: "; $P --scene stress --width 3840 --height 2160 --spp-per-step 4 2>/dev/null | python -c "$S"
done
unset ARE_B200_LIB
echo -n "stress-1M host SAH: "; $P --scene stress --width 3840 --height 2160 --spp-per-step 4 --builder 0 2>/dev/null | python -c "$S"
M=l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_st.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
ncu --metrics $M --clock-control none -k regex:k_render_path -c 1 python bench.py --scene stress --width 3840 --height 2160 --spp-per-step 2 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic 2>&1 | grep -E "l1tex|lts|smsp|gpu__|dram" > gpurun_out/r02x_stress_l1_quant.txt
cat gpurun_out/r02x_stress_l1_quant.txt
