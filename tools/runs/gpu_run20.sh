P="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic"
S='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"],1), "Msamples/s", round(d["ms_per_step"],2), "ms", round(d["mrays_per_s"],1), "Mrays/s", r["kernel"][:40])'
for v in default sc1 sc2 b10 b8 sc1b10 sc1b8; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "stress-1M $v: "; $P --scene stress --width 3840 --height 2160 --spp-per-step 4 2>/dev/null | python -c "$S"
done
for v in default sc1 sc2; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  echo -n "rtiow $v: "; $P --scene rtiow_final --width 1200 --height 675 --spp-per-step 100 2>/dev/null | python -c "$S"
  echo -n "stress-1M bvh4 $v: "; $P --scene stress --width 3840 --height 2160 --spp-per-step 4 --traversal 4 2>/dev/null | python -c "$S"
done
M=l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_local.sum,l1tex__data_pipe_lsu_wavefronts_mem_global.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum
for v in default sc1 sc2; do
  if [ "$v" = default ]; then unset ARE_B200_LIB; else export ARE_B200_LIB=$PWD/variants/libare_b200_$v.so; fi
  ncu --metrics $M --clock-control none -k regex:k_render_path -c 1 python bench.py --scene stress --width 3840 --height 2160 --spp-per-step 2 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-configs --no-strong --no-traffic 2>&1 | grep -E "l1tex|lts|smsp|gpu__" > gpurun_out/r02w_stress_l1_$v.txt
  echo "== stress $v"; cat gpurun_out/r02w_stress_l1_$v.txt
done
