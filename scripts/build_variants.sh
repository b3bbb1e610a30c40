#!/bin/bash
# builds A/B variants of the library into warp_b200/lib/variants/<name>.so; usage: build_variants.sh name="-DFLAG=1 ..." ...
set -e
cd "$(dirname "$0")/.."
mkdir -p warp_b200/lib/variants
for spec in "$@"; do
    name="${spec%%=*}"; flags="${spec#*=}"
    WARP_B200_NVCC_EXTRA="$flags" python -m warp_b200.build --force > /dev/null
    cp warp_b200/lib/libwarp_b200.so "warp_b200/lib/variants/$name.so"
    echo "built $name ($flags)"
done
python -m warp_b200.build --force > /dev/null
