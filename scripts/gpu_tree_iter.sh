#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -q -x --timeout 120 -p no:cacheprovider 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_query.py tests/test_gpu_bvh_query.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
bash scripts/gpu_launches.sh
AB_SIZES="810 2237 7072" bash scripts/gpu_ab_build.sh
