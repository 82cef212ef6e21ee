#!/bin/bash
# ncu --set full capture of the tracking kernel (exact and fma variants), 1 GPU.
OUT=gpurun_out/${1:-ncu}
mkdir -p $OUT
for mode in "" "--fma"; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 2 -c 1 \
    -o $OUT/prof_track${mode} -f python bench.py --quick --steps 1 --warmup 1 --turns 3 --no-cpu-baseline $mode > $OUT/ncu${mode}.log 2>&1
done
ls -la $OUT
