#!/bin/bash
# GPU-box session: variant sweep (quick bench, exact + fma), parity tests, one ncu capture.
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bash scripts/sweep_variants.sh $TAG > $OUT/sweep.txt 2>&1
cat $OUT/sweep.txt
timeout 1200 python -m pytest tests -m gpu -q -s --durations=12 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -E "passed|failed|exit" $OUT/pytest_gpu.log | tail -4
timeout 500 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 2 -c 1 \
    -o $OUT/prof_track -f python bench.py --quick --steps 1 --warmup 1 --turns 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT | tail -5
