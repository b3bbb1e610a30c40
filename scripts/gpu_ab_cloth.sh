#!/bin/bash
# A/B of library variants on the headline and the C4 cloth frame
run() { name=$1; shift; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['extra']; print('$name: %.1f Mq/s  e2e %.1f  signed %.1f M  rays %.2f G  cloth frame %.2f ms' % (d['value']/1e6, d['e2e']['value']/1e6, e['signed']['queries_per_s']/1e6, e['rays']['rays_per_s']/1e9, e['cloth']['frame_ms']))"; }
for rep in 1 2; do
run default
for f in warp_b200/lib/variants/*.so; do run "$(basename $f .so)" WARP_B200_LIB=$PWD/$f; done
done
