set -x
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.max,smsp__cycles_active.avg,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum"
IFS=";" read -ra CFGS <<< "${K2_CFGS:-1 1 1;1 0 1;1 1 3;1 0 3;2 1 1;2 0 1;2 1 3;2 0 3}"; unset IFS
for cfg in "${CFGS[@]}"; do
  set -- $cfg
  kind=random; [ "$1" = 2 ] && kind=trained; [ -n "$4" ] && kind=$4
  VSOM_TC_TIER=$1 VSOM_TC_PAIR=$2 VSOM_TC_STAGGER=$3 timeout 300 ncu --metrics $M --clock-control none -k regex:score_tc_kernel -s 1 -c 1 --csv --log-file gpurun_out/k2m_t$1_p$2_s$3$4.csv python tests/profile_k2.py 524288 $kind > gpurun_out/k2m_t$1_p$2_s$3$4.log 2>&1
done
