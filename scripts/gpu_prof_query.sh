#!/bin/bash
mkdir -p gpurun_out
export PROF_NQ=$((1<<22))
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_query_point|k_query_ray|k_refit' -c 6 -f -o gpurun_out/prof_query_r01 python scripts/prof_driver.py > gpurun_out/prof_query.log 2>&1
echo "exit $?"; tail -2 gpurun_out/prof_query.log
