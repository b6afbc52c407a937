#!/bin/bash
# k_resolve phase clocks for prebuilt library variants (variants/*.so built with -DFT_RS_CLOCK), L2 flushed and not.
# Usage: tools/rs_variants.sh v0_clock v1_clock ...
LIB=fasttrack_b200/_build/libfasttrack_b200.so
cp $LIB /tmp/lib_orig.so
for v in "$@"; do
  cp variants/$v.so $LIB
  for fl in 1 0; do
    echo "== $v flush=$fl"
    RS_FLUSH=$fl timeout 120 python tools/rs_clock.py 2>&1 | grep "^cta 0\|^cta 7\|rror"
  done
done
cp /tmp/lib_orig.so $LIB
