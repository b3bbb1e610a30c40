#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_query_ray' -s 2 -c 1 -f -o gpurun_out/r01_query_ray python scripts/prof_ray_driver.py > gpurun_out/prof_ray.log 2>&1
echo "exit $?"; tail -2 gpurun_out/prof_ray.log
ncu -i gpurun_out/r01_query_ray.ncu-rep --page raw --csv > gpurun_out/r01_query_ray.raw.csv 2>/dev/null
