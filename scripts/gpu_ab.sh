#!/bin/bash
run() { name=$1; shift; env "$@" python bench.py --steps 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name: %.1f Mq/s  %.2f ms/step  e2e %.1f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"; }
timeout 600 python -m pytest tests/test_gpu_query.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
for q in 0 4 8 12 16 24; do run quorum$q WARP_B200_LEAF_QUORUM=$q; done
