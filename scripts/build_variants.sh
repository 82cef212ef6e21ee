#!/bin/bash
# Builds tuning variants of libxtb200 (see xtrack_b200/build.py: XTB_LIB_SUFFIX / XTB_EXTRA_DEFINES).
# usage: scripts/build_variants.sh name1 "defs1" name2 "defs2" ...
rm -f xtrack_b200/libxtb200_*.so
while [ $# -ge 2 ]; do
  XTB_LIB_SUFFIX=_$1 XTB_EXTRA_DEFINES="$2" python -m xtrack_b200.build | tail -1 &
  shift 2
done
wait
