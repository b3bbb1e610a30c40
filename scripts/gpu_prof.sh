#!/bin/bash
mkdir -p gpurun_out
export PROF_NQ=$((1<<22))
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python scripts/prof_driver.py > gpurun_out/prof_launches.log 2>&1
echo "launch list exit $?"
export PROF_NQ=$((1<<20))
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_query_point|k_query_ray|k_hierarchy|k_onesweep|k_refit|k_morton|k_scene' -s 8 -c 14 -f -o gpurun_out/prof_r01 python scripts/prof_driver.py > gpurun_out/prof_full.log 2>&1
echo "full exit $?"
tail -3 gpurun_out/prof_full.log
ls -la gpurun_out/
