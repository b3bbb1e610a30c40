#!/bin/bash
# ncu --set full of the merge kernels (build + refit) at C2 size
mkdir -p gpurun_out
export PROF_NQ=$((1<<16))
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_refit_wave' -s 1 -c 1 -f -o gpurun_out/prof_merge python scripts/prof_driver.py > gpurun_out/prof_merge.log 2>&1
echo "exit $?"
