#!/bin/bash
# first GPU contact: parity tests per file (separate processes), smoke, short bench, reference dumps
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_build test_gpu_query; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -30 gpurun_out/$f.log
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 900 python baseline/ref_cuda.py golden > gpurun_out/ref_golden.log 2>&1; echo "ref golden exit $?" >> gpurun_out/summary.txt
timeout 1200 python baseline/ref_cuda.py timing > gpurun_out/ref_timing.log 2>&1; echo "ref timing exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/smoke.log; tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err; tail -5 gpurun_out/ref_golden.log; tail -5 gpurun_out/ref_timing.log
