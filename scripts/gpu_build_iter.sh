#!/bin/bash
# build/refit iteration: parity tests for build + query, then launch list
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -q -x --timeout 120 -p no:cacheprovider 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_query.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4
bash scripts/gpu_launches.sh
