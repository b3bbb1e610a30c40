#!/bin/bash
export PROF_NQ=$((1<<22)) PROF_SIGN=1
timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_query_point<\(bool\)1|k_query_ray' -c 2 --csv python scripts/prof_driver.py 2>/dev/null | grep -E "k_query" | awk -F'","' '{print substr($5,1,40), $(NF-2), $(NF-1), $NF}'
