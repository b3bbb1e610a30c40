#!/bin/bash
mkdir -p gpurun_out
export PROF_NQ=$((1<<18))
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_onesweep_pass|k_merge|k_leaves|k_morton' -s 10 -c 8 -f -o gpurun_out/prof_build_r01 python scripts/prof_driver.py > gpurun_out/prof_build.log 2>&1
echo "exit $?"
