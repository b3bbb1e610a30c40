#!/bin/bash
# A/B of library variants on the bench workload (query kernels): headline + signed + rays
run() { name=$1; shift; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['extra']; print('$name: %.1f Mq/s  %.2f ms/step  e2e %.1f  signed %.1f M  rays %.2f G' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, e['signed']['queries_per_s']/1e6, e['rays']['rays_per_s']/1e9))"; }
for rep in 1 2; do
run default
for f in warp_b200/lib/variants/*.so; do run "$(basename $f .so)" WARP_B200_LIB=$PWD/$f; done
done
