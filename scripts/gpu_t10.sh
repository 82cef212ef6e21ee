#!/bin/bash
# GPU-box session: ncu captures of the radiation kernels + GPU parity tests.
TAG=${1:-t10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for wl in clic_dr_quantum lep_quantum; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 1 -c 1 \
    -o $OUT/prof_$wl -f python bench.py --workload $wl --particles 100000 --quick --steps 1 --warmup 1 --turns 1 --no-cpu-baseline > $OUT/ncu_$wl.log 2>&1
done
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
ls -la $OUT
