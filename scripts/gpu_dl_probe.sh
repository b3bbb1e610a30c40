#!/bin/bash
export PROF_NQ=$((1<<22))
for q in 0 8; do
  WARP_B200_LEAF_QUORUM=$q timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum --clock-control none -k regex:k_query_point -s 1 -c 1 --csv python scripts/prof_driver.py 2>/dev/null | grep -E "k_query_point" | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | sed "s/^/q=$q /"
done
