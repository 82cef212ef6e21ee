#!/bin/bash
# GPU-box session: variant sweep (quick bench, exact + fma) and one ncu --set full capture
# (with source counters) of the default tracking kernel.
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bash scripts/sweep_variants.sh $TAG > $OUT/sweep.txt 2>&1
cat $OUT/sweep.txt
for mode in "" "--fma"; do
timeout 500 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 2 -c 1 \
    -o $OUT/prof_track${mode} -f python bench.py --quick --steps 1 --warmup 1 --turns 3 --no-cpu-baseline $mode > $OUT/ncu_full${mode}.log 2>&1
done
ls -la $OUT
