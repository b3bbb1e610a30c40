#!/bin/bash
# refit / build iteration: parity tests, C2 launch list, then build/refit timing across sizes for the refit variants
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -q -x --timeout 120 -p no:cacheprovider 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_query.py tests/test_gpu_capture.py tests/test_gpu_bvh_query.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
bash scripts/gpu_launches.sh
echo "== split wavefront refit (default)"; timeout 600 python scripts/scale_probe.py 810 2237 4473 7072
echo "== monolithic wavefront"; WARP_B200_REFIT=wave timeout 600 python scripts/scale_probe.py 810 7072
echo "== atomic refit";    WARP_B200_REFIT=atomic timeout 600 python scripts/scale_probe.py 810 2237 7072
