#!/bin/bash
# ncu counters of the per-action kernels at one saturated iteration of a TestEm3 pass, for the
# library given as $1 (path or "default") -> gpurun_out/layout_$2.csv
# usage: scratch/ncu_layout.sh <lib|default> <tag>
lib=$1; tag=$2
[ "$lib" != default ] && export CELERITAS_B200_LIB=$lib
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
M=$M,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
M=$M,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum
M=$M,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum
M=$M,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
M=$M,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread
ncu --metrics $M --clock-control none \
    -k regex:'k_pre_step|k_along_step|k_discrete_select|k_interact_lists|k_post_tail|k_initialize_tracks|k_end_pass' \
    --launch-skip 900 --launch-count 11 --csv --log-file gpurun_out/layout_$tag.csv \
    python scratch/one_pass.py > gpurun_out/layout_$tag.log 2>&1
python scratch/ncu_layout_table.py gpurun_out/layout_$tag.csv $tag
