#!/bin/bash
# A/B of library variants on rebuild time at a few sizes
run() { name=$1; shift; echo "== $name"; env "$@" timeout 300 python scripts/scale_probe.py ${AB_SIZES:-810 1415 2237 4473}; }
for rep in 1 2; do
run default
for f in warp_b200/lib/variants/*.so; do [ -f "$f" ] && run "$(basename $f .so)" WARP_B200_LIB=$PWD/$f; done
done
