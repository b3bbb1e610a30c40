#!/bin/bash
# A/B of library variants on build / refit time: every warp_b200/lib/variants/*.so plus the default build
run() { name=$1; shift; echo "== $name"; env "$@" timeout 300 python scripts/scale_probe.py ${AB_SIZES:-810 2237}; }
for rep in 1 2; do
run default
for f in warp_b200/lib/variants/*.so; do run "$(basename $f .so)" WARP_B200_LIB=$PWD/$f; done
done
