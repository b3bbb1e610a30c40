#!/bin/bash
# round profiles: launch list of the bench command + ncu --set full of the dominant kernels
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --cloth-frames 20 --skip refcuda,parity ${BENCH_NCU_ARGS:-} > gpurun_out/${R}_bench_under_ncu.json 2> gpurun_out/${R}_bench_under_ncu.err
echo "launch list exit $?"
# the query kernel at full bench size (one launch), build + refit kernels at C2 size
export PROF_NQ=$((1<<24))
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_query_point|k_unpack_results' -s 2 -c 2 -f -o gpurun_out/${R}_query_point python scripts/prof_driver.py > gpurun_out/${R}_prof_q.log 2>&1
echo "query exit $?"
export PROF_NQ=$((1<<20))
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_scene|k_morton|k_onesweep|k_leaves|k_merge|k_refit|k_deep|k_query_ray' -s 12 -c 16 -f -o gpurun_out/${R}_build_refit python scripts/prof_driver.py > gpurun_out/${R}_prof_b.log 2>&1
echo "build exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_refit|k_plan' -s 0 -c 12 -f -o gpurun_out/${R}_refit_wave python scripts/prof_refit_driver.py > gpurun_out/${R}_prof_r.log 2>&1
echo "refit-wave exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_query_ray' -s 2 -c 1 -f -o gpurun_out/${R}_query_ray python scripts/prof_ray_driver.py > gpurun_out/${R}_prof_ray.log 2>&1
echo "ray exit $?"
for f in ${R}_query_point ${R}_build_refit ${R}_refit_wave ${R}_query_ray; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
