#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_query.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "order or bit_exact or c3" 2>&1 | tail -3
bash scripts/gpu_launches.sh 2>&1 | grep -i "onesweep\|merge\|leaves"
timeout 600 python scripts/scale_probe.py 810 2237 7072
timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench: %.1f Mq/s  %.2f ms/step  e2e %.1f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"
