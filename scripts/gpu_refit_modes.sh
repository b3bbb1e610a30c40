#!/bin/bash
for bits in 30 63; do
  for mode in split atomic; do
    echo "== bits $bits mode $mode"; PROBE_BITS=$bits WARP_B200_REFIT=$mode timeout 600 python scripts/scale_probe.py 1415 4473 7072 2>&1 | grep -v "^$"
  done
done
