#!/bin/bash
# quick iteration: GPU parity tests (both files) + short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit $?"; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
    r = d['roofline']; e = d.get('extra', {})
    print('value %.1f Mq/s  ms/step %.2f  e2e %.1f Mq/s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))
    print('pairs/q %.1f tris/q %.1f  roofline frac %.3f' % (r['pair_fetches_per_query'], r['tri_fetches_per_query'], r['frac']))
    for k in ('build_ms_constructor','build_ms_kernels','refit_ms'):
        print(k, e.get(k))
    print('rays', e.get('rays'))
    print('clocks', d['clocks'], 'cpu', d.get('cpu_baseline',{}).get('value'))
except Exception as ex:
    print('parse failed', ex)
PY
