#!/bin/bash
for c in $((1<<21)) $((3<<20)) $((1<<22)) $((6<<20)) $((1<<23)); do
  echo -n "chunk $c: "; WARP_B200_HOST_CHUNK=$c python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e %.1f Mq/s' % (d['e2e']['value']/1e6))"
done
