#!/bin/bash
# A/B of library variants on build / refit time (+ the bench headline, whose query ordering shares the build kernels)
run() { name=$1; shift; echo "== $name"; env "$@" timeout 300 python scripts/scale_probe.py ${AB_SIZES:-810 2237};
  env "$@" python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   bench %.1f Mq/s  %.2f ms/step  e2e %.1f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"; }
for rep in 1 2; do
run default
for f in warp_b200/lib/variants/*.so; do [ -f "$f" ] && run "$(basename $f .so)" WARP_B200_LIB=$PWD/$f; done
done
