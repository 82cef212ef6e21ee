#!/bin/bash
# Round-2 session 14 (thin radiation run loop): radiation -- tests (Philox, golden moments), benches of the four radiation
# workloads, by-function profile of the CLIC-DR mean-model kernel.
TAG=${1:-r02s14}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_radiation.py tests/test_philox.py -m gpu -q -s > $OUT/pytest_rad.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_rad.log
grep -E "passed|failed|FAILED|Error|sigma_delta" $OUT/pytest_rad.log | tail -20
for wl in clic_dr_mean clic_dr_quantum lep_mean lep_quantum; do
  timeout 400 python bench.py --workload $wl --quick --steps 2 --warmup 1 --turns 2 --particles 300000 > $OUT/bench_$wl.json 2>> $OUT/bench.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_$wl.json')); print('$wl', '%.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'])
except Exception as e: print('$wl FAILED', e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xtb_track_kernel -s 1 -c 1 \
    -o $OUT/prof_clic_mean -f python bench.py --workload clic_dr_mean --particles 300000 --quick --steps 1 --warmup 1 --turns 1 --no-cpu-baseline > $OUT/ncu_clic_mean.log 2>&1
