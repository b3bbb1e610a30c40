#!/bin/bash
# refit / build iteration: parity tests, C2 launch list, then build/refit timing across sizes
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -q -x --timeout 120 -p no:cacheprovider 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_query.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "refit or cloth" 2>&1 | tail -3
bash scripts/gpu_launches.sh
timeout 600 python scripts/scale_probe.py 810 2237 4473 7072
