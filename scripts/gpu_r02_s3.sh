#!/bin/bash
# Round-2 session 3: ncu --set full of the thin kernel at 125 000 particles (NPT 1 and 3) and 10^6.
TAG=${1:-r02s3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for cfg in "125000 1" "125000 3" "1000000 3"; do
  set -- $cfg
  XTB_NPT_FORCE=$2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 2 -c 1 \
    -o $OUT/prof_n$1_npt$2 -f python bench.py --quick --steps 1 --warmup 1 --turns 3 --no-cpu-baseline --particles $1 > $OUT/ncu_n$1_npt$2.log 2>&1
done
ls -la $OUT
